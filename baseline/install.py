"""Install the UNMODIFIED reference under ``baseline/_ref`` (git-ignored; travels to the GPU box with the snapshot).

    python -m baseline.install [--source /root/reference] [--force]

The reference is a directory of scripts and packages (``model/ data/ core/ config/ tests/ split.py infer.py ...``) with
no ``setup.py`` / ``pyproject.toml``, so ``pip install --target baseline/_ref /root/reference`` has nothing to build
("does not appear to be a Python project"; the attempt is recorded in ``baseline/_ref/INSTALL.json``).  The install is
therefore a verbatim copy of the Python sources and JSON configs: nothing is edited, and a SHA-256 manifest of every
copied file is written next to them so that a test can prove it.  Consumers:
  * ``bench.py --impl reference`` - times the reference's own modules on the host cores;
  * ``tests/test_boundary.py``    - drives the reference's ``core/logger.parse`` / ``split.get_datasets`` /
    ``tests/test_tiling_setup.py`` against this package's ``model`` / ``data.tile_stitcher`` / ``predtiler`` replacements.
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
ITEMS = ("model", "data", "core", "config", "tests", "split.py", "infer.py", "eval.py", "sample.py", "requirement.txt", "LICENSE")
SKIP_DIRS = {"__pycache__", ".git"}


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def manifest(root):
    out = {}
    for dp, dn, fn in os.walk(root):
        dn[:] = [d for d in dn if d not in SKIP_DIRS]
        for f in sorted(fn):
            if f in ("INSTALL.json",) or f.endswith(".pyc"):
                continue
            p = os.path.join(dp, f)
            out[os.path.relpath(p, root)] = _sha(p)
    return out


def install(source="/root/reference", force=False):
    """Returns the install record (dict).  No-op when ``baseline/_ref`` is already there and ``source`` is absent."""
    rec_path = os.path.join(DEST, "INSTALL.json")
    if os.path.exists(rec_path) and not force:
        return json.load(open(rec_path))
    if not os.path.isdir(source):
        raise FileNotFoundError(f"reference not found at {source} and {DEST} is not installed")
    pip = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                          "--target", os.path.join(HERE, "_pip_probe"), source], capture_output=True, text=True)
    shutil.rmtree(os.path.join(HERE, "_pip_probe"), ignore_errors=True)
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    for it in ITEMS:
        src = os.path.join(source, it)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(DEST, it), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.exists(src):
            shutil.copy2(src, os.path.join(DEST, it))
    rec = {"source": source, "method": "verbatim copy (the reference has no build metadata for pip)",
           "pip_returncode": pip.returncode, "pip_tail": (pip.stderr or pip.stdout).strip().splitlines()[-1:] if (pip.stderr or pip.stdout) else [],
           "files": manifest(DEST)}
    with open(rec_path, "w") as fh:
        json.dump(rec, fh, indent=0, sort_keys=True)
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    r = install(a.source, a.force)
    print(f"baseline/_ref: {len(r['files'])} files from {r['source']} ({r['method']}); pip rc={r['pip_returncode']} {r['pip_tail']}")
