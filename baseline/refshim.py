"""Import the unmodified reference from ``baseline/_ref`` (see ``baseline/install.py``).

``activate()`` puts ``baseline/_ref`` at the front of ``sys.path`` and stubs the third-party modules the reference imports
at module level but which this image does not have (``albumentations``, ``skimage.io``, ``tensorboardX``, ``lmdb``): none
of them is touched by the sampling / tiling path.  Use it in a process of its own (the reference's top-level package
names are ``model``, ``data`` and ``core``)."""
import importlib
import os
import sys
import types

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "INSTALL.json"))


def activate():
    if not available():
        raise FileNotFoundError("baseline/_ref is not installed: run `python -m baseline.install` where /root/reference exists")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("albumentations", "skimage", "skimage.io", "tensorboardX", "lmdb"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["skimage.io"], "imread"):
        sys.modules["skimage.io"].imread = None
        sys.modules["skimage"].io = sys.modules["skimage.io"]
    if not hasattr(sys.modules["albumentations"], "Compose"):       # constructed (never applied) by SplitDataset(enable_transforms=True)
        sys.modules["albumentations"].Compose = lambda *a, **k: None
        sys.modules["albumentations"].HorizontalFlip = lambda *a, **k: None
    if not hasattr(sys.modules["tensorboardX"], "SummaryWriter"):
        sys.modules["tensorboardX"].SummaryWriter = None
    return REF


def build_sampler(which, unet_kw, diffusion_kw):
    """Reference sampler + UNet built DIRECTLY from the reference classes (``define_G`` raises TypeError for sr3 / ddpm,
    model/networks.py:159-170).  ``which``: 'sr3' | 'ddpm' | 'indi'."""
    activate()
    if which == "sr3":
        from model.sr3_modules.diffusion import GaussianDiffusion
        from model.sr3_modules.unet import UNet
        net = UNet(**unet_kw)
        return GaussianDiffusion(net, **diffusion_kw), net
    from model.ddpm_modules.unet import UNet
    net = UNet(**unet_kw)
    if which == "ddpm":
        from model.ddpm_modules.diffusion import GaussianDiffusion
        return GaussianDiffusion(net, **diffusion_kw), net
    from model.ddpm_modules.indi import InDI
    return InDI(net, **diffusion_kw), net
