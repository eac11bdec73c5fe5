#!/usr/bin/env python
"""Throughput of the text boundary of call_mods (host code; runs with or without a GPU):
native feature-file parser / call-line formatter of libdsp_b200 vs the per-line Python of the
reference (oracle port, a bounded sample).  One JSON line per stage.

    python tools/bench_textio.py [--sites 65536] [--threads N]"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepsignal_plant_b200 import feature_io, synthetic  # noqa: E402
from oracle import features_oracle, callmods_oracle  # noqa: E402  (CPU baseline only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=65536)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--cpu_sample", type=int, default=4096)
    a = ap.parse_args()
    n = a.sites
    feats = synthetic.make_features(min(n, 8192), 13, 16, seed=1)
    info = synthetic.make_sampleinfo(min(n, 8192), seed=1)
    base = [feature_io.features_to_str(info[i], feats["kmer"][i], feats["base_means"][i].astype(np.float64).round(6),
                                       feats["base_stds"][i].astype(np.float64).round(6), feats["base_signal_lens"][i],
                                       feats["signals"][i].astype(np.float64).round(6), 0) for i in range(len(info))]
    lines = (base * ((n + len(base) - 1) // len(base)))[:n]
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "features.tsv")
        with open(path, "w") as f:
            f.write("\n".join(lines) + "\n")
        size = os.path.getsize(path)
        best = 1e30
        for rep in range(3):
            t0 = time.perf_counter()
            batches = 0
            last = None
            for b in feature_io.FeatureFileReader(path, 13, 16, batch_sites=65536, pinned=False, nthreads=a.threads):
                batches += 1
                last = b
            best = min(best, time.perf_counter() - t0)
        t0 = time.perf_counter()
        features_oracle.read_features(lines[:a.cpu_sample])
        t_ref = time.perf_counter() - t0
        print(json.dumps({"stage": "parse feature file -> float32 batches (dsp_parse_features)", "sites": n, "bytes": size,
                          "threads": a.threads, "sites_per_s": n / best, "MB_per_s": size / best / 1e6,
                          "reference_python_sites_per_s": a.cpu_sample / t_ref,
                          "reference_sample": "%d lines, per-line Python of call_modifications.py:55-127 (oracle port), 1 core" % a.cpu_sample}))
        probs = np.random.default_rng(0).random((last.n, 2)).astype(np.float32)
        labels = probs.argmax(1).astype(np.int32)
        best = 1e30
        for rep in range(3):
            t0 = time.perf_counter()
            out = feature_io.format_calls(last, probs, labels, nthreads=a.threads)
            best = min(best, time.perf_counter() - t0)
        m = min(a.cpu_sample, last.n)
        sinfo = last.sampleinfo()[:m]
        t0 = time.perf_counter()
        callmods_oracle.call_lines(sinfo, last.kmer[:m].numpy(), probs[:m])
        t_ref = time.perf_counter() - t0
        print(json.dumps({"stage": "probabilities -> call_mods lines (dsp_format_calls)", "sites": last.n, "bytes": len(out),
                          "threads": a.threads, "sites_per_s": last.n / best, "MB_per_s": len(out) / best / 1e6,
                          "reference_python_sites_per_s": m / t_ref,
                          "reference_sample": "%d sites, per-site loop of call_modifications.py:175-188 (oracle port), 1 core" % m}))


if __name__ == "__main__":
    main()
