"""The drop-in claim of SURVEY 8(b), exercised with the UNMODIFIED reference front end (baseline/_ref, installed by
``python -m baseline.install``): its ``core/logger.parse`` on its own config JSONs, its ``split.py`` (imported as is), its
``tests/test_tiling_setup.py`` - with ``model``, ``data.tile_stitcher``, ``data.tiling_manager`` and ``predtiler`` resolved to
this package by ``diffsplitting_b200.install()``.  Each case runs ``tests/boundary_driver.py`` in a process of its own."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

from baseline import install as INST
from baseline import refshim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not refshim.available(), reason="baseline/_ref not installed (python -m baseline.install)")
CONFIGS = ["splitting_cifar10.json", "splitting_hagen_indi_single_ch.json", "splitting_hagen_indi_joint.json",
           "sr_sr3_16_128.json", "sr_sr3_64_512.json", "splitting.json"]


def _drive(*args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "boundary_driver.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@needs_ref
def test_installed_reference_is_unmodified():
    rec = json.load(open(os.path.join(refshim.REF, "INSTALL.json")))
    assert len(rec["files"]) >= 40 and "split.py" in rec["files"] and "model/sr3_modules/unet.py" in rec["files"]
    assert INST.manifest(refshim.REF) == rec["files"]
    src = rec["source"]
    if os.path.isdir(src):                      # in the build container: byte-identical to the reference checkout
        for rel, digest in rec["files"].items():
            with open(os.path.join(src, rel), "rb") as fh:
                assert hashlib.sha256(fh.read()).hexdigest() == digest, rel


@needs_ref
def test_reference_front_end_resolves_to_this_package_and_tiles_490():
    out = _drive("host")
    assert out["model_is_ours"] and out["predtiler_is_ours"] and out["stitcher_is_ours"]
    assert out["split_file"] == os.path.join("baseline", "_ref", "split.py")
    assert out["tiled_len"] == 490                                    # notebooks/EvaluateJointIndi.ipynb:1683
    assert out["patch_location_0"] == [0, 0, 0] and out["patch_location_last"] == [9, 1536, 1536]
    assert out["item_shapes"] == {"input": [1, 512, 512], "target": [2, 512, 512]}
    assert {k: v[0] for k, v in out["configs_parsed"].items()} == {
        "splitting_cifar10.json": "ddpm", "splitting_hagen_indi_single_ch.json": "indi", "splitting_hagen_indi_joint.json": "joint_indi",
        "sr_sr3_16_128.json": "sr3", "sr_sr3_64_512.json": "sr3", "splitting.json": "sr3"}


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS)
def test_create_model_test_visuals_through_the_reference_front_end(cfg):
    out = _drive("gpu", cfg)
    assert out["netG"].startswith("diffsplitting_b200.model.samplers.")
    vis = out["visuals"]
    assert set(vis) == {"prediction", "input", "target"}
    for k, (shape, dtype, device, finite) in vis.items():
        assert dtype == "torch.float32" and device == "cpu" and finite, (k, vis[k])      # model/model.py:102-112
    if cfg == "splitting_hagen_indi_joint.json":
        assert out["reference_tiling_test"] == "passed"               # the reference's tests/test_tiling_setup.py
        assert out["stitch_490_shape"] == [10, 2048, 2048, 2] and out["stitch_490_max_abs_err"] == 0.0
