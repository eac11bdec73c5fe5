"""ORACLE (test infrastructure, not product code): run the UNMODIFIED reference's ``_call_mods``
(``call_modifications.py:130-192``) with the reference's own ``ModelBiLSTM`` on the CPU and write its lines.

    CUDA_VISIBLE_DEVICES="" python oracle/run_ref_callmods.py --out lines.tsv --n 3000 --batch 512 \
        --weight_seed 1234 --feature_seed 21 --rng_seed 77 [--threads N] [--time]

Meant to run in a subprocess with ``CUDA_VISIBLE_DEVICES=""``: the reference decides CPU vs CUDA once, at
import (``utils/constants_torch.py:6``).  Inputs are the synthetic features of
``deepsignal_plant_b200/synthetic.py`` turned into the nested Python lists the reference's reader hands
over (``call_modifications.py:55-127``); weights are the reference constructor's own under
``torch.manual_seed(weight_seed)``; ``torch.manual_seed(rng_seed)`` right before the call fixes the
``torch.randn`` initial states (``models.py:169-176``).  With ``--time`` it prints one JSON line with the
wall-clock seconds of the ``_call_mods`` call (list -> tensor + forward + per-site text loop), which is the
CPU baseline SURVEY.md 8(d) asks for.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def features_batch(n, seq_len, signal_len, feature_seed):
    import numpy as np
    from deepsignal_plant_b200 import synthetic
    feats = synthetic.make_features(n, seq_len, signal_len, seed=feature_seed)
    info = synthetic.make_sampleinfo(n, seed=feature_seed)
    return (info, feats["kmer"].astype(np.int64).tolist(), feats["base_means"].tolist(), feats["base_stds"].tolist(),
            feats["base_signal_lens"].astype(np.int64).tolist(), feats["signals"].tolist(), [0] * n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--n", type=int, default=3000)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--weight_seed", type=int, default=1234)
    ap.add_argument("--feature_seed", type=int, default=21)
    ap.add_argument("--rng_seed", type=int, default=77)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    import torch
    from oracle import ref_import
    ref_models, ref_cm = ref_import.import_reference("models", "call_modifications")
    assert not ref_cm.use_cuda, "run with CUDA_VISIBLE_DEVICES=\"\": this script is the reference's CPU path"
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    torch.manual_seed(a.weight_seed)
    model = ref_models.ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True, module="both_bilstm")
    model.eval()
    batch = features_batch(a.n, 13, 16, a.feature_seed)
    best, times = None, []
    for _ in range(a.repeat):
        torch.manual_seed(a.rng_seed)
        t0 = time.perf_counter()
        lines, acc, nb = ref_cm._call_mods(batch, model, a.batch, 0)
        dt = time.perf_counter() - t0
        times.append(dt)
        best = dt if best is None else min(best, dt)
    if a.out:
        with open(a.out, "w") as f:
            f.write("\n".join(lines) + "\n")
    if a.time:
        print(json.dumps({"sites": a.n, "seconds": best, "sites_per_s": a.n / best, "threads": torch.get_num_threads(),
                          "batch": a.batch, "batches": nb, "all_seconds": times, "reference": ref_import.ref_root()}))


if __name__ == "__main__":
    main()
