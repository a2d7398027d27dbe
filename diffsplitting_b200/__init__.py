"""diffsplitting_b200 - B200-native (sm_100a) sampling hot path of rayanirban/DiffSplitting.

    import diffsplitting_b200 as dsb
    dsb.install()            # makes `import model`, `import data.tile_stitcher`, `import predtiler.dataset`
                             # resolve to this package, so split.py / infer.py / eval.py run unchanged
    netG = dsb.model.networks.define_G(opt)

See include/diffsplit_b200.h for the C ABI and INTEGRATION.md for how it plugs into the reference.
"""
import sys
import types

from . import _lib  # noqa: F401
from . import core, data, model  # noqa: F401

__version__ = "0.1.0"


def install(override_existing=True):
    """Alias the reference's module names to this package (the duck-typed boundary of SURVEY.md section 8b)."""
    from .data import tile_stitcher, tiled_pred, tiling_manager
    from .model import base_model, model as model_mod, networks, samplers, time_predictor, unet

    def put(name, mod):
        if override_existing or name not in sys.modules:
            sys.modules[name] = mod

    put("model", model)
    put("model.model", model_mod)
    put("model.networks", networks)
    put("model.base_model", base_model)
    dm = types.ModuleType("model.ddpm_modules")                   # only the inference-side classes live here
    dm.time_predictor = time_predictor
    put("model.ddpm_modules", dm)
    put("model.ddpm_modules.time_predictor", time_predictor)
    # data.* : only the tiling modules are replaced; the TIFF/LMDB dataset classes stay the reference's own
    put("data.tiling_manager", tiling_manager)
    put("data.tile_stitcher", tile_stitcher)
    pt = types.ModuleType("predtiler")
    ptd = types.ModuleType("predtiler.dataset")
    ptd.get_tile_manager = tiled_pred.get_tile_manager
    ptd.get_tiling_dataset = tiled_pred.get_tiling_dataset
    pt.dataset = ptd
    put("predtiler", pt)
    put("predtiler.dataset", ptd)
    return {"samplers": samplers, "unet": unet}
