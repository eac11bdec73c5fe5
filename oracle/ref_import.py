"""ORACLE (test infrastructure, not product code): import the unmodified reference package.

Looks in ``oracle/_ref`` (built by ``oracle/build_ref.py``; travels to the GPU box) and, in the build
container, falls back to ``/root/reference``.  ``deepsignal_plant.call_modifications`` imports
``extract_features`` which imports ``h5py`` and ``statsmodels`` at module scope
(``extract_features.py:13,24``); neither is in this image and only the fast5 path uses them, so empty
stand-ins are registered for exactly those names when they are missing.

``deepsignal_plant.utils.constants_torch.use_cuda`` is evaluated at import (``constants_torch.py:6``):
import with ``CUDA_VISIBLE_DEVICES=""`` (see ``cpu_env``) to get the reference's CPU path on a GPU box.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "deepsignal_plant", "models.py")) or \
        os.path.exists("/root/reference/deepsignal_plant/models.py")


def ref_root():
    if os.path.exists(os.path.join(REF_DIR, "deepsignal_plant", "models.py")):
        return REF_DIR
    if os.path.exists("/root/reference/deepsignal_plant/models.py"):
        return "/root/reference"
    raise ImportError("the reference is not installed: run `python oracle/build_ref.py` where /root/reference exists")


def _stub(name, attrs=()):
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
        return
    except ImportError:
        pass
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        sub = ".".join(parts[:i])
        if sub not in sys.modules:
            mod = types.ModuleType(sub)
            mod.__dsp_stub__ = True
            sys.modules[sub] = mod
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], mod)
    for a in attrs:
        setattr(sys.modules[name], a, None)


def import_reference(*modules):
    """import_reference("models", "call_modifications") -> the reference's modules, in order."""
    root = ref_root()
    if root not in sys.path:
        sys.path.insert(0, root)
    _stub("h5py")
    _stub("statsmodels")
    _stub("statsmodels.robust", ("mad",))
    out = [importlib.import_module("deepsignal_plant." + m) for m in modules]
    return out[0] if len(out) == 1 else out


def cpu_env():
    """Environment for a subprocess in which the reference must take its CPU path."""
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    return env
