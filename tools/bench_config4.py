#!/usr/bin/env python
"""BASELINE.json configs[4] end to end on the device: ``call_mods`` -> ``call_freq`` over N per-read calls, one
process per GPU, nothing becomes text in between.

Every rank classifies its share of synthetic sites with ModelBiLSTM (batches of 65 536, in-kernel Philox states),
turns each batch's probabilities into the record columns the reference's text round trip would yield
(``chain.records_from_probs``), keeps them in HBM, and the ranks then aggregate all calls per site with the NVLink
exchange (``dsp_freq_aggregate_distributed``).  One JSON line on rank 0: both phases timed on the device, whole-job
rates, and the integer invariants of the table (every callable call counted exactly once, slices ordered).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
        tools/bench_config4.py --calls 1000000000
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepsignal_plant_b200 import chain, freq_dist as fd, synthetic  # noqa: E402
from deepsignal_plant_b200.models import ModelBiLSTM  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calls", type=int, default=100_000_000, help="per-read calls (= classified sites) over all ranks")
    ap.add_argument("--coverage", type=int, default=20)
    ap.add_argument("--prob_cf", type=float, default=0.0,
                    help="random-init weights put prob_1 at 0.5 +- 0.003: with the CLI default 0.5 nothing would be callable")
    ap.add_argument("--batch", type=int, default=65536)
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ndev = torch.cuda.device_count()
    device = local % ndev
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    if world > 1:
        dist.init_process_group("nccl" if ndev >= world else "gloo", **({"device_id": dev} if ndev >= world else {}))
        grp = fd.TorchGroup()
    else:
        grp = fd.SoloGroup()
    per = a.calls // world
    lo = rank * per
    bounds = np.array([per * r for r in range(world + 1)], np.uint64)
    n_sites = max(per * world // a.coverage, 1)
    torch.manual_seed(1234)
    model = ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True, precision="fp16", max_batch=a.batch, seed=rank).cuda(device).eval()
    base = synthetic.make_features(a.batch, 13, 16, seed=rank)
    names = ("kmer", "base_means", "base_stds", "base_signal_lens", "signals")
    pool = [tuple(torch.from_numpy(np.ascontiguousarray(np.roll(base[k], 977 * b, axis=0))).to(dev) for k in names) for b in range(4)]
    calls = chain.DeviceCalls(per, dev)
    for i in range(3):
        model(*pool[i % 4])
    torch.cuda.synchronize()
    grp.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    done = 0
    b = 0
    while done < per:
        m = min(a.batch, per - done)
        _, probs = model(*(t[:m] for t in pool[b % 4]))
        calls.append(fd.synth_keys(lo + done, lo + done + m, n_sites, dev), probs, model.last_labels)
        done += m
        b += 1
    e1.record()
    torch.cuda.synchronize()
    classify_ms = e0.elapsed_time(e1)
    key, p0, p1, lab = calls.columns()
    win = (int(per * 1.3) + (1 << 16)) * 32
    be = fd.DeviceBackend(rank, world, device, win, grp.all_gather_object)
    be.aggregate_tensors(key[:4096], p0[:4096], p1[:4096], lab[:4096], lo, bounds, a.prob_cf)     # first call: allocations, CUB set-up
    grp.barrier()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    rows, n_call = be.aggregate_tensors(key, p0, p1, lab, lo, bounds, a.prob_cf)
    freq_s = time.perf_counter() - t1
    total_s = time.perf_counter() - t0
    stages = be.timing()
    mine = rows.cpu().numpy().reshape(-1).view(fd.SITE_ROW)
    callable_local = int((~((p0 - p1).abs() < a.prob_cf)).sum())
    info = grp.all_gather_object((classify_ms, freq_s, total_s, len(mine), int(mine["cov"].sum()), n_call, callable_local,
                                  bool(len(mine) == 0 or (np.diff(mine["first"].astype(np.int64)) > 0).all()),
                                  int(mine["met"].sum()), int((lab == 1).sum())))
    be.close()
    if rank == 0:
        cls = max(x[0] for x in info) * 1e-3
        frq = max(x[1] for x in info)
        tot = max(x[2] for x in info)
        n = per * world
        print(json.dumps({
            "metric": "configs[4]: call_mods + call_freq over %d per-read calls on %d GPU(s), device resident" % (n, world),
            "calls": n, "world": world, "sites": sum(x[3] for x in info), "prob_cf": a.prob_cf,
            "classify_s": cls, "classified_sites_per_s": n / cls, "call_freq_s": frq, "call_freq_records_per_s": n / frq,
            "end_to_end_s": tot, "end_to_end_calls_per_s": n / tot, "call_freq_stage_ms_rank0": stages,
            "every_callable_call_counted_once": sum(x[4] for x in info) == sum(x[5] for x in info) == sum(x[6] for x in info),
            "met_equals_label1_calls": (sum(x[8] for x in info) == sum(x[9] for x in info)) if a.prob_cf == 0.0 else None,
            "slices_ordered": all(x[7] for x in info)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
