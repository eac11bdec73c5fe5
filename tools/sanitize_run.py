#!/usr/bin/env python
"""One forward of the tcgen05 path under compute-sanitizer (racecheck / synccheck / memcheck):
>= 2 waves of CTA pairs (default 40 960 sites = 320 tiles -> 640 layer_kernel CTAs on 148 SMs),
FP16 operands, in-kernel Philox states.  DSP_B200_BRANCH_DUAL=0|1 (read at handle creation) picks
layer_kernel or branch_kernel for the hidden-128 layers; run it once each.

    compute-sanitizer --tool racecheck python tools/sanitize_run.py --sites 40960
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepsignal_plant_b200 import synthetic  # noqa: E402
from deepsignal_plant_b200.models import ModelBiLSTM  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sites", type=int, default=40960)
ap.add_argument("--passes", type=int, default=1)
a = ap.parse_args()
torch.manual_seed(1234)
model = ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True, precision="fp16", max_batch=a.sites).cuda(0).eval()
f = synthetic.make_features(a.sites, 13, 16, seed=0)
args = [torch.from_numpy(f[k]).cuda(0) for k in ("kmer", "base_means", "base_stds", "base_signal_lens", "signals")]
for _ in range(a.passes):
    probs = model(*args)[1]
torch.cuda.synchronize()
p = probs.cpu().numpy()
assert np.isfinite(p).all()
print("sanitize_run: %d sites, %d launches, dual=%s, mean p1 %.4f" % (a.sites, model.launch_count(),
      os.environ.get("DSP_B200_BRANCH_DUAL", "auto"), float(p[:, 1].mean())))
