"""Device-side counterparts of the reference's ``core`` helpers that sit next to the sampling path."""
from . import psnr  # noqa: F401
