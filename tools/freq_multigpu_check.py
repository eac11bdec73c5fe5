#!/usr/bin/env python
"""Multi-GPU call_freq check (launch under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/freq_multigpu_check.py --records 8000000

Every rank takes a contiguous shard of the same seeded synthetic call records (file order =
generation order), the shards are aggregated with one key-hash all-to-all over NCCL
(call_mods_freq.aggregate_records_distributed) and rank 0 compares the merged table, bit for
bit, with the single-GPU aggregation of all records.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsignal_plant_b200 import call_mods_freq as cf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=4_000_000)
    ap.add_argument("--prob_cf", type=float, default=0.2)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # the same seeded records on every rank, generated on the device; rank r keeps shard r (file order)
    n = args.records
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    chrom = torch.randint(0, 5, (n,), device=dev, generator=g)
    pos = torch.randint(0, max(n // 20, 1), (n,), device=dev, generator=g)      # coverage ~20 per site over 5 chromosomes
    key = (chrom << cf.POS_BITS) | pos
    p1 = torch.round(torch.rand(n, device=dev, generator=g, dtype=torch.float64) * 1e6) / 1e6
    p0 = torch.round((1.0 - p1) * 1e6) / 1e6
    label = (p1 > p0).to(torch.int32)
    lo, hi = rank * n // world, (rank + 1) * n // world
    gidx = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    for it in range(3):                                   # first passes warm NCCL / CUB up
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = cf.aggregate_tensors_distributed(key[lo:hi], p0[lo:hi], p1[lo:hi], label[lo:hi], gidx, args.prob_cf, sort_by_key=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        k, first, s0, s1, met, unmet, cov = cf._aggregate_tensors(key, p0, p1, label, args.prob_cf, True, dev)
        want = torch.stack([k, first, s0.view(torch.int64), s1.view(torch.int64), met.long(), unmet.long(), cov.long()], 1)
        ok = res.shape == want.shape and bool((res == want).all())
        print(json.dumps({"check": "freq multi-GPU == single-GPU, bit for bit", "ok": bool(ok), "world": world,
                          "records": n, "sites": int(res.shape[0]), "seconds": float(tt[0]),
                          "records_per_s": n / float(tt[0])}), flush=True)
        if not ok:
            sys.exit(1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
