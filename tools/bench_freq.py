#!/usr/bin/env python
"""Throughput of the call_freq aggregation (dsp_freq_aggregate) with records resident in HBM,
against the HBM roofline, with the reference's per-record Python loop (oracle port, one core,
bounded sample) beside it.  One JSON line.

    python tools/bench_freq.py [--records 50000000] [--coverage 20]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepsignal_plant_b200 import call_mods_freq as cf  # noqa: E402
from oracle import freq_oracle  # noqa: E402  (CPU baseline only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=50_000_000)
    ap.add_argument("--coverage", type=int, default=20)
    ap.add_argument("--prob_cf", type=float, default=0.5)
    ap.add_argument("--cpu_sample", type=int, default=300_000)
    a = ap.parse_args()
    n = a.records
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    # 5 chromosomes x n/coverage/5 positions, global record order = generation order (BASELINE.json configs[4])
    npos = max(n // a.coverage // 5, 1)
    chrom = torch.randint(0, 5, (n,), device=dev, generator=g)
    pos = torch.randint(0, npos, (n,), device=dev, generator=g)
    key = (chrom << cf.POS_BITS) | pos
    p1 = torch.round(torch.rand(n, device=dev, generator=g, dtype=torch.float64) * 1e6) / 1e6
    p0 = torch.round((1.0 - p1) * 1e6) / 1e6
    lab = (p1 > p0).to(torch.int32)
    for _ in range(2):
        out = cf._aggregate_tensors(key, p0, p1, lab, a.prob_cf, False, dev)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = cf._aggregate_tensors(key, p0, p1, lab, a.prob_cf, False, dev)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = min(ts)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    alg_bytes = 28 * n                      # key 8 + p0 8 + p1 8 + label 4, each record read once
    # CPU baseline: the reference's loop (call_mods_freq.py:29-74) restated, on a sample of the same records
    m = min(a.cpu_sample, n)
    k_h, p0_h, p1_h, l_h = key[:m].cpu().numpy(), p0[:m].cpu().numpy(), p1[:m].cpu().numpy(), lab[:m].cpu().numpy()
    lines = ["chr%d\t%d\t+\t%d\tr\tt\t%r\t%r\t%d\tAACGT" % (int(k) >> cf.POS_BITS, int(k) & ((1 << cf.POS_BITS) - 1),
                                                            int(k) & 0xffff, float(x), float(y), int(l))
             for k, x, y, l in zip(k_h.tolist(), p0_h.tolist(), p1_h.tolist(), l_h.tolist())]
    t0 = time.perf_counter()
    freq_oracle.aggregate(lines, a.prob_cf)
    t_cpu = time.perf_counter() - t0
    print(json.dumps({"metric": "call_freq records/s (dsp_freq_aggregate, records resident in HBM)", "records": n,
                      "sites": int(out[0].shape[0]), "seconds": t, "value": n / t, "unit": "records/s",
                      "roofline": {"bound": "hbm", "achieved": alg_bytes / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": alg_bytes / t / 1e9 / peaks["hbm_gbs"],
                                   "algorithmic_bytes_per_record": 28},
                      "cpu_baseline": {"value": m / t_cpu, "unit": "records/s", "cores": 1, "kind": "port",
                                       "sample": "%d records incl. line parsing, oracle/freq_oracle.py" % m}}))


if __name__ == "__main__":
    main()
