#!/usr/bin/env python
"""Multi-GPU call_freq throughput (BASELINE.json configs[4]): records resident in HBM, sharded over the ranks,
exchanged over NVLink peer memory, aggregated, rows returned to their home ranks.  One JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
        tools/bench_freq_dist.py --records_per_rank 50000000 [--no_check] [--iters 5]
    python tools/bench_freq_dist.py --records_per_rank 50000000        # one GPU, same code path (world = 1)
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepsignal_plant_b200 import freq_dist as fd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records_per_rank", type=int, default=50_000_000)
    ap.add_argument("--coverage", type=int, default=20)
    ap.add_argument("--prob_cf", type=float, default=0.5)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no_check", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ndev = torch.cuda.device_count()
    device = local % ndev
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl" if ndev >= world else "gloo", **({"device_id": torch.device("cuda", device)} if ndev >= world else {}))
        grp = fd.TorchGroup()
    else:
        grp = fd.SoloGroup()
    out = fd.measure(grp, device, a.records_per_rank, a.coverage, a.prob_cf, a.iters, check=not a.no_check)
    if grp.rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
        gbs = 28 * out["records"] / out["seconds"] / 1e9 / world
        out.update({"metric": "call_freq records/s, %d GPU(s), records resident in HBM" % world, "unit": "records/s",
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s per GPU", "frac": gbs / peaks["hbm_gbs"],
                                 "algorithmic_bytes_per_record": 28}})
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
