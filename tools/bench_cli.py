#!/usr/bin/env python
"""End-to-end `call_mods` command line on a large synthetic feature file: text -> native parser ->
pinned batches -> CUDA forward -> native formatter -> output file.  Prints one JSON line.

    python tools/bench_cli.py [--sites 1000000]
    python tools/bench_cli.py --archive [--reads 2000]     # decoded-reads archive -> extract + call in one pass"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepsignal_plant_b200 import cli, feature_io, synthetic  # noqa: E402
from deepsignal_plant_b200.models import ModelBiLSTM  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=1_000_000)
    ap.add_argument("--nproc", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--archive", action="store_true", help="input is a decoded-reads .npz (extract_features.save_reads)")
    ap.add_argument("--reads", type=int, default=2000)
    ap.add_argument("--extract", action="store_true", help="time `extract` (archive -> the reference's feature file) instead")
    a = ap.parse_args()
    if a.archive or a.extract:
        return archive(a)
    base_n = 8192
    feats = synthetic.make_features(base_n, 13, 16, seed=1)
    info = synthetic.make_sampleinfo(base_n, seed=1)
    lines = [feature_io.features_to_str(info[i], feats["kmer"][i], feats["base_means"][i].astype(np.float64).round(6),
                                        feats["base_stds"][i].astype(np.float64).round(6), feats["base_signal_lens"][i],
                                        feats["signals"][i].astype(np.float64).round(6), 0) for i in range(base_n)]
    block = ("\n".join(lines) + "\n").encode()
    with tempfile.TemporaryDirectory() as tmp:
        path, ckpt, out = os.path.join(tmp, "features.tsv"), os.path.join(tmp, "m.ckpt"), os.path.join(tmp, "calls.tsv")
        reps = max(1, a.sites // base_n)
        with open(path, "wb") as f:
            for _ in range(reps):
                f.write(block)
        n = reps * base_n
        torch.manual_seed(1234)
        torch.save(ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True).state_dict(), ckpt)
        argv = ["call_mods", "-i", path, "-m", ckpt, "-o", out, "--host_threads", str(a.nproc)]
        cli.main(argv)                                  # first run: page cache, CUDA context
        t0 = time.perf_counter()
        cli.main(argv)
        dt = time.perf_counter() - t0
        nout = sum(1 for _ in open(out, "rb"))
        print(json.dumps({"metric": "call_mods command line, feature file -> calls file (sites/s, wall clock incl. model load)",
                          "sites": n, "lines_written": nout, "seconds": dt, "value": n / dt, "unit": "sites/s",
                          "input_bytes": os.path.getsize(path), "output_bytes": os.path.getsize(out), "host_threads": a.nproc}))


def archive(a):
    from deepsignal_plant_b200 import extract_features as ef
    base = synthetic.make_reads(min(a.reads, 200), seed=1, mean_bases=8000, long_every=9)
    reads = [dict(base[i % len(base)], readname="r%06d" % i) for i in range(a.reads)]
    with tempfile.TemporaryDirectory() as tmp:
        path, ckpt, out = os.path.join(tmp, "reads.npz"), os.path.join(tmp, "m.ckpt"), os.path.join(tmp, "calls.tsv")
        ef.save_reads(path, reads)
        torch.manual_seed(1234)
        torch.save(ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True).state_dict(), ckpt)
        argv = ["call_mods", "-i", path, "-m", ckpt, "-o", out, "--motifs", "CG", "--f5_batch_size", "130"]
        if a.extract:
            argv = ["extract", "-i", path, "-o", out, "--motifs", "CG", "--f5_batch_size", "130", "--host_threads", str(a.nproc)]
        cli.main(argv)
        t0 = time.perf_counter()
        cli.main(argv)
        dt = time.perf_counter() - t0
        nout = sum(1 for _ in open(out, "rb"))
        what = ("extract command line, decoded-reads archive -> feature file (sites/s, wall clock)" if a.extract else
                "call_mods command line, decoded-reads archive -> calls file (sites/s, wall clock incl. archive and model load)")
        print(json.dumps({"metric": what,
                          "reads": a.reads, "sites": nout, "seconds": dt, "value": nout / dt, "unit": "sites/s",
                          "input_bytes": os.path.getsize(path), "output_bytes": os.path.getsize(out)}))


if __name__ == "__main__":
    main()
