"""Throughput of dsp_extract_features on synthetic decoded reads.

    python tools/bench_extract.py [--reads 2000] [--bases 8000] [--motifs CG] [--steps 20]

Times, with CUDA events on the launching stream and inputs resident in HBM: the whole call, and the
per-read normalisation alone (a call with zero sites).  Algorithmic bytes: 2 B per raw sample read once
+ 24 B per event touched + 1 040 B per site written (13x16: 4*13*4 + 13*16*4).  The oracle
(oracle/extract_oracle.py, numpy, one core) is timed next to it on a bounded sample of the same reads.
Prints one JSON line."""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsignal_plant_b200 import extract_features as ef, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2000)
    ap.add_argument("--bases", type=int, default=8000)
    ap.add_argument("--motifs", default="CG")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-reads", type=int, default=60)
    a = ap.parse_args()
    K, S = 13, 16
    t0 = time.time()
    base = synthetic.make_reads(min(a.reads, 200), seed=1, mean_bases=a.bases, long_every=9)
    reads = [dict(base[i % len(base)], readname="r%06d" % i) for i in range(a.reads)]
    batch = ef.pack_reads(reads)
    motif_seqs = ef.get_motif_seqs(a.motifs)
    t1 = time.time()
    sites = ef.find_sites(batch, motif_seqs, 0, None, K)
    t_find = time.time() - t1
    n_samples, n_events, n_sites = int(batch.raw.shape[0]), int(batch.ev_len.shape[0]), len(sites)
    none = ef.Sites(sites.site_read[:0], sites.site_ev[:0], sites.pos[:0], sites.pos_in_strand[:0])
    batch.to_device(torch.device("cuda", 0))
    torch.cuda.synchronize()

    def timed(s):
        out = None
        for _ in range(a.warmup):
            out = ef.extract_tensors(batch, s, K, S, seed=1, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(a.steps):
            ef.extract_tensors(batch, s, K, S, seed=i, out=out)     # everything resident, outputs reused: launches only
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    ms_all, ms_norm = timed(sites), timed(none)
    ms_site = ms_all - ms_norm
    bytes_norm = 2 * n_samples
    bytes_site = n_sites * (K * 24 + 1040) + 2 * int(batch.ev_len[(sites.site_ev[:, None] + np.arange(-6, 7)[None, :])].sum())
    cpu = reads[:a.cpu_reads]
    from oracle import extract_oracle as eo
    t1 = time.time()
    feats, _ = eo.extract_features(cpu, "mad", motif_seqs, 0, None, K, S, 1, rng=random.Random(0))
    t_cpu = time.time() - t1
    print(json.dumps({
        "metric": "feature extraction from decoded reads (mad normalisation + 13x16 features)", "unit": "sites/s",
        "value": n_sites / (ms_all * 1e-3), "samples_per_s": n_samples / (ms_all * 1e-3),
        "reads": a.reads, "samples": n_samples, "events": n_events, "sites": n_sites, "motifs": a.motifs,
        "ms": {"call": ms_all, "read_scale_kernel": ms_norm, "site_features_kernel": ms_site,
               "host_find_sites": t_find * 1e3},
        "roofline": {"bound": "hbm", "unit": "GB/s",
                     "read_scale_kernel": {"algorithmic_bytes": bytes_norm, "achieved": bytes_norm / (ms_norm * 1e-3) / 1e9},
                     "site_features_kernel": {"algorithmic_bytes": bytes_site, "achieved": bytes_site / (ms_site * 1e-3) / 1e9}},
        "cpu_baseline": {"value": len(feats) / t_cpu, "unit": "sites/s", "cores": 1, "kind": "port",
                         "sample": "%d reads, %d sites, %.1f s (numpy oracle of _extract_features)" % (len(cpu), len(feats), t_cpu)},
        "setup_s": t1 - t0}))


if __name__ == "__main__":
    main()
