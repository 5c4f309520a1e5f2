"""Install the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` -- TEST / BENCH INFRASTRUCTURE.

The reference has no setup.py / pyproject.toml (a script directory plus ``spikingjelly.zip``), so
``pip install /root/reference`` does not apply.  This recipe is the equivalent: it copies the Python sources of the
path (``snn_model/``, ``metric/pytorch_ssim``, ``metric/Fid_score.py``, ``metric/IS_score.py``) byte for byte and
extracts the vendored SpikingJelly zip into a directory named ``spikingjelly`` (the zip has no top-level package
directory).  Nothing is edited.  ``baseline/_ref`` is listed in .gitignore (reference sources never enter the history)
but not in .gpurunignore, so it travels to the GPU box, where ``bench.py --impl reference`` and the metric parity tests
import the reference's own code from it.  ``__graft_entry__.build()`` runs this whenever /root/reference is mounted.

    python oracle/install_ref.py        # prints the destination, or why nothing was installed
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/Spiking-Diffusion-release"
DST = os.path.join(ROOT, "baseline", "_ref")
COPY = ["snn_model/__init__.py", "snn_model/vae_model.py", "snn_model/vq_diffusion.py", "snn_model/snn_layers.py",
        "metric/pytorch_ssim/__init__.py", "metric/Fid_score.py", "metric/IS_score.py"]


def installed() -> bool:
    return os.path.isfile(os.path.join(DST, "INSTALLED.json"))


def install(force: bool = False):
    """Returns the install directory, or None when the reference is not mounted (an earlier install is kept)."""
    if not os.path.isfile(os.path.join(SRC, "spikingjelly.zip")):
        return DST if installed() else None
    if installed() and not force:
        return DST
    rel = os.path.join(DST, "Spiking-Diffusion-release")
    manifest = {}
    for f in COPY:
        src = os.path.join(SRC, f)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(rel, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[f] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    pkg = os.path.join(DST, "spikingjelly")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    os.makedirs(pkg)
    zipfile.ZipFile(os.path.join(SRC, "spikingjelly.zip")).extractall(pkg)
    manifest["spikingjelly.zip"] = hashlib.sha256(open(os.path.join(SRC, "spikingjelly.zip"), "rb").read()).hexdigest()
    with open(os.path.join(DST, "INSTALLED.json"), "w") as fh:
        json.dump({"source": SRC, "note": "unmodified copies; sha256 of each source file", "files": manifest}, fh, indent=1)
    return DST


if __name__ == "__main__":
    d = install(force="--force" in sys.argv)
    print(d if d else f"reference not mounted at {SRC} and no earlier install under {DST}")
