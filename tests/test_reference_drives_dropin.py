"""The drop-in claim, executed: the UNMODIFIED reference's own ``_call_mods`` (call_modifications.py:130-192)
and worker loop ``_call_mods_q`` (:195-259) drive ``deepsignal_plant_b200.ModelBiLSTM``, and their output lines
are compared with the reference's CPU run at equal seeds (the committed fixture, and a fresh CPU run of the
reference in a subprocess).  The reference comes from ``oracle/_ref`` (``oracle/build_ref.py``, travels to the GPU
box) or, in the build container, ``/root/reference``."""
import os
import queue
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

import cases
from oracle import ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
needs_ref = pytest.mark.skipif(not ref_import.available(), reason="oracle/_ref is not built (python oracle/build_ref.py)")
PROB_TOL = 1e-3
LABEL_AGREEMENT = 0.9999


def compare_lines(lines, gold, prob_tol):
    assert len(lines) == len(gold)
    flips = 0
    worst = 0.0
    for a, b in zip(lines, gold):
        wa, wb = a.split("\t"), b.split("\t")
        assert len(wa) == 10 and wa[:6] == wb[:6] and wa[9] == wb[9]            # sample info and 5-mer: identical
        worst = max(worst, abs(float(wa[6]) - float(wb[6])), abs(float(wa[7]) - float(wb[7])))
        flips += wa[8] != wb[8]
    assert worst <= prob_tol, worst
    return flips, worst


def test_ref_import_is_the_unmodified_reference():
    if not ref_import.available():
        pytest.skip("oracle/_ref is not built")
    m = ref_import.import_reference("models")
    assert m.ModelBiLSTM.__module__ == "deepsignal_plant.models"
    assert os.path.realpath(m.__file__).startswith((os.path.realpath(os.path.join(ROOT, "oracle", "_ref")), "/root/reference"))
    if os.path.isdir("/root/reference") and os.path.isdir(os.path.join(ROOT, "oracle", "_ref")):
        for f in ("models.py", "call_modifications.py", "call_mods_freq.py", "utils/txt_formater.py"):
            assert open(os.path.join(ROOT, "oracle", "_ref", "deepsignal_plant", f), "rb").read() == \
                open(os.path.join("/root/reference/deepsignal_plant", f), "rb").read(), f


@needs_ref
def test_reference_cpu_run_reproduces_the_committed_fixture(tmp_path):
    # the fixture tests/golden/callmods_77.tsv.gz came from this very script's recipe; re-running the reference on
    # this box's CPU (other thread count, other oneDNN dispatch) must give the same lines up to float32 noise
    e = cases.MANIFEST["callmods"]
    out = str(tmp_path / "ref_lines.tsv")
    subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_ref_callmods.py"), "--out", out, "--n", str(e["n"]),
                    "--batch", str(e["batch"]), "--weight_seed", str(e["weight_seed"]), "--feature_seed", str(e["feature_seed"]),
                    "--rng_seed", str(e["rng_seed"])], check=True, env=ref_import.cpu_env(), timeout=600)
    gold = cases.read_gz("callmods_%d.tsv.gz" % e["rng_seed"]).splitlines()
    flips, worst = compare_lines(open(out).read().splitlines(), gold, 5e-6)
    assert flips <= 1


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_reference_call_mods_drives_the_dropin_model(precision):
    from deepsignal_plant_b200.models import ModelBiLSTM
    from run_ref_callmods import features_batch
    ref_cm = ref_import.import_reference("call_modifications")
    assert ref_cm.use_cuda                                   # on the GPU box the reference hands over CUDA tensors
    e = cases.MANIFEST["callmods"]
    torch.manual_seed(e["weight_seed"])
    model = ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True, module="both_bilstm", precision=precision,
                        state_mode="init_hidden").cuda(0).eval()
    batch = features_batch(e["n"], 13, 16, e["feature_seed"])
    torch.manual_seed(e["rng_seed"])
    lines, acc, nb = ref_cm._call_mods(batch, model, e["batch"], 0)      # the reference's code, our module
    assert model.launch_count() > 0 and nb == (e["n"] + e["batch"] - 1) // e["batch"]
    gold = cases.read_gz("callmods_%d.tsv.gz" % e["rng_seed"]).splitlines()
    flips, worst = compare_lines(lines, gold, 2e-6 if precision == "fp32" else PROB_TOL)
    print("reference _call_mods + drop-in (%s): %d lines, max |dprob| %.1e, %d label flips" % (precision, len(lines), worst, flips))
    assert flips <= max(1, int(len(gold) * (1 - LABEL_AGREEMENT)))


@pytest.mark.gpu
@needs_ref
def test_reference_worker_loop_builds_and_drives_the_dropin(tmp_path):
    # _call_mods_q (call_modifications.py:195-259): constructor call with the reference's positional arguments,
    # torch.load(map_location=cpu) -> model_dict.update -> load_state_dict -> .cuda(device) -> .eval(), queue loop
    # until "kill".  Only the class behind the name ModelBiLSTM is substituted; a fresh CPU run of the reference
    # itself (other seeds than the fixture) is the expectation.
    from deepsignal_plant_b200.models import ModelBiLSTM as DropIn
    from run_ref_callmods import features_batch
    ref_cm, ref_models = ref_import.import_reference("call_modifications", "models")
    n, bs, wseed, fseed, rseed = 1536, 512, 5, 31, 99
    torch.manual_seed(wseed)
    ckpt = str(tmp_path / "both_bilstm.b13_s16_epoch0.ckpt")
    torch.save(ref_models.ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True).state_dict(), ckpt)   # train.py:161
    out = str(tmp_path / "ref_lines.tsv")
    subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_ref_callmods.py"), "--out", out, "--n", str(n),
                    "--batch", str(bs), "--weight_seed", str(wseed), "--feature_seed", str(fseed), "--rng_seed", str(rseed)],
                   check=True, env=ref_import.cpu_env(), timeout=600)
    gold = open(out).read().splitlines()

    class InitHiddenModel(DropIn):
        def __init__(self, *a, **kw):
            super().__init__(*a, state_mode="init_hidden", **kw)
    args = types.SimpleNamespace(seq_len=13, signal_len=16, layernum1=3, layernum2=1, class_num=2, dropout_rate=0, hid_rnn=256,
                                 n_vocab=16, n_embed=4, is_base="yes", is_signallen="yes", model_type="both_bilstm", batch_size=bs)
    fq, pq = queue.Queue(), queue.Queue()
    fq.put(features_batch(n, 13, 16, fseed))
    fq.put("kill")
    saved, orig = ref_cm.ModelBiLSTM, ref_cm._call_mods
    ref_cm.ModelBiLSTM = InitHiddenModel
    try:
        # the constructor draws from the generator too, so seed where the CPU run seeds: right before _call_mods
        def seeded(batch, model, batch_size, device=0):
            torch.manual_seed(rseed)
            return orig(batch, model, batch_size, device)
        ref_cm._call_mods = seeded
        ref_cm._call_mods_q(ckpt, fq, pq, None, args, 0)
    finally:
        ref_cm.ModelBiLSTM = saved
        ref_cm._call_mods = orig
    lines = pq.get_nowait()
    flips, worst = compare_lines(lines, gold, PROB_TOL)
    print("reference _call_mods_q + drop-in: %d lines, max |dprob| %.1e, %d label flips" % (len(lines), worst, flips))
    assert flips <= 1
