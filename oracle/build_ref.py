"""ORACLE (test infrastructure, not product code): install the UNMODIFIED reference into
``oracle/_ref`` so that it travels to the GPU box with the snapshot.

    python oracle/build_ref.py [--force]

``oracle/_ref/`` is git-ignored (the reference's sources never enter this repository's history) but not
gpurun-ignored.  The recipe is the base contract's offline install: the tree under ``/root/reference`` is
copied to a scratch directory (the install writes ``build/`` and ``*.egg-info`` next to ``setup.py`` and
``/root/reference`` is read-only) and installed with

    python -m pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy>

``--no-deps``: its pins (torch<=1.11, numpy<1.20, h5py, statsmodels ...) cannot be met here; it runs on
this image's torch / numpy, and ``oracle/ref_import.py`` stubs the two absent modules that only the fast5
path uses.  Nothing under ``deepsignal_plant_b200/`` imports from here: only tests, smoke() and bench.py's
CPU-baseline / ``--impl reference`` legs do, as the checker / the baseline.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("DSP_REFERENCE_SRC", "/root/reference")
TARGET = os.path.join(HERE, "_ref")


def installed():
    return os.path.exists(os.path.join(TARGET, "deepsignal_plant", "models.py"))


def build(force=False):
    """Returns the install directory, or None when /root/reference is not present (the GPU box)."""
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(REF_SRC, "deepsignal_plant")):
        return None
    tmp = tempfile.mkdtemp(prefix="dsp_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not installed():
            raise RuntimeError("reference install failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-2000:]))
        # the stand-alone scripts are not part of the wheel; combine_call_mods_freq_files.py is the one the
        # freq tests cross-check against
        os.makedirs(os.path.join(TARGET, "scripts"), exist_ok=True)
        for f in os.listdir(os.path.join(REF_SRC, "scripts")):
            if f.endswith(".py"):
                shutil.copy2(os.path.join(REF_SRC, "scripts", f), os.path.join(TARGET, "scripts", f))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
