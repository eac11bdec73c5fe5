"""CPU: the C-ABI library builds, loads and exports every symbol include/dsp_b200.h declares;
without a GPU the product path fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import cases
from deepsignal_plant_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = torch.cuda.is_available()


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    assert _native.lib().dsp_abi_version() == 1


def test_exports_match_header():
    header = open(os.path.join(ROOT, "include", "dsp_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dsp_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_native.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dsp_[a-z_0-9]+)", out))
    assert declared <= exported


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_shipped_library_has_no_result_changing_switch():
    # measurement hooks that change results (skipping activation stores) exist only in variant builds
    # (-DDSP_MEAS_SKIP_Y); the shipped .so must not even contain the switch's name
    blob = open(_native.LIB_PATH, "rb").read()
    for needle in (b"MEAS_SKIP", b"SKIP_Y_NOW"):
        assert needle not in blob, needle


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    L = _native.lib()
    cfg = _native.DspConfig(seq_len=13, signal_len=16, num_layers1=3, num_layers2=1, num_classes=2, hidden_size=256,
                            vocab_size=16, embedding_size=4, is_base=1, is_signallen=1, module=0, device=0,
                            precision=0, reserved=0, max_batch=64)
    h = C.c_void_p()
    rc = L.dsp_create(C.byref(h), C.byref(cfg))
    assert rc != 0 and not h.value
    assert b"no CPU path" in L.dsp_last_error()


def test_create_rejects_bad_config():
    L = _native.lib()
    cfg = _native.DspConfig(seq_len=13, signal_len=16, num_layers1=3, num_layers2=1, num_classes=2, hidden_size=256,
                            vocab_size=16, embedding_size=4, is_base=1, is_signallen=1, module=7, device=0,
                            precision=0, reserved=0, max_batch=64)
    h = C.c_void_p()
    assert L.dsp_create(C.byref(h), C.byref(cfg)) == 1
    assert b"--model_type is not right!" in L.dsp_last_error()
    with pytest.raises(_native.DspError):
        _native.check(1, "x")


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_model_forward_has_no_cpu_fallback():
    m = cases.build_model(cases.MANIFEST["forward"]["both_small_odd"])
    x = torch.zeros(3, 5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, x, x, x, torch.zeros(3, 5, 8))
    with pytest.raises(RuntimeError, match="no CPU"):
        m.forward_host(np.zeros((3, 5)), np.zeros((3, 5)), np.zeros((3, 5)), np.zeros((3, 5)), np.zeros((3, 5, 8)))
