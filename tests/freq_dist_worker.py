"""One rank of the multi-rank call_freq tests (launched by torchrun from test_freq_dist.py)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from deepsignal_plant_b200 import call_mods_freq as cf  # noqa: E402
from deepsignal_plant_b200 import freq_dist as fd  # noqa: E402


def synth_records(n, dev, seed=7, coverage=20):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    chrom = torch.randint(0, 5, (n,), device=dev, generator=g)
    pos = torch.randint(0, max(n // coverage // 5, 1), (n,), device=dev, generator=g)
    key = (chrom << cf.POS_BITS) | pos
    p1 = torch.round(torch.rand(n, device=dev, generator=g, dtype=torch.float64) * 1e6) / 1e6
    p0 = torch.round((1.0 - p1) * 1e6) / 1e6
    return key, p0, p1, (p1 > p0).to(torch.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="device", choices=["device", "standin"])
    ap.add_argument("-i", "--input_path", action="append", nargs="+", default=[])
    ap.add_argument("-o", "--result_file", default=None)
    ap.add_argument("--prob_cf", type=float, default=0.0)
    ap.add_argument("--sort", action="store_true")
    ap.add_argument("--bed", action="store_true")
    ap.add_argument("--gzip", action="store_true")
    ap.add_argument("--contigs", default=None)
    ap.add_argument("--tensor_check", action="store_true")
    ap.add_argument("--records", type=int, default=1000000)
    ap.add_argument("--window_records", type=int, default=0)
    ap.add_argument("--expect_overflow", action="store_true")
    ap.add_argument("--home_rows", action="store_true", help="row_placement 1: rows go to the rank that parsed their first record")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    grp = fd.TorchGroup()
    if a.backend == "standin":
        from dist_standin import StandInBackend
        backend = StandInBackend(grp)
        device = None
    else:
        device = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
        torch.cuda.set_device(device)
        backend = None
    if a.tensor_check:
        dev = torch.device("cuda", device)
        n = a.records
        key, p0, p1, lab = synth_records(n, dev)                  # the same records on every rank; rank r keeps shard r
        bounds = np.array([n * r // world for r in range(world + 1)], np.uint64)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        win = (a.window_records or int(n / world * 1.3) + 65536) * 32
        be = fd.DeviceBackend(rank, world, device, win, grp.all_gather_object)
        try:
            rows, n_call = be.aggregate_tensors(key[lo:hi], p0[lo:hi], p1[lo:hi], lab[lo:hi], lo, bounds, a.prob_cf,
                                                balanced=not a.home_rows)
            overflow = 0
        except Exception as e:
            if not (a.expect_overflow and "overflow" in str(e)):
                raise
            overflow, rows = 1, None
        if a.expect_overflow:
            tot = sum(grp.all_gather_object(overflow))
            if rank == 0:
                print(json.dumps({"ok": tot == world, "overflow_ranks": tot}), flush=True)
            be.close()
            dist.destroy_process_group()
            return
        mine = rows.cpu().numpy().reshape(-1).view(fd.SITE_ROW)
        every = grp.all_gather_object(mine)
        rb = be.last_row_bounds
        in_range = bool(len(mine) == 0 or (mine["first"].min() >= rb[rank] and (rank == world - 1 or mine["first"].max() < rb[rank + 1])))
        in_range = all(grp.all_gather_object(in_range))
        if rank == 0:
            got = np.concatenate(every)                          # rank order = first-appearance order
            k, first, s0, s1, met, unmet, cov = cf._aggregate_tensors(key, p0, p1, lab, a.prob_cf, False, dev)
            ok = (len(got) == k.shape[0] and (got["key"].view(np.int64) == k.cpu().numpy()).all()
                  and (got["first"].astype(np.int64) == first.cpu().numpy()).all()
                  and (got["s0"].view(np.int64) == s0.cpu().numpy().view(np.int64)).all()
                  and (got["s1"].view(np.int64) == s1.cpu().numpy().view(np.int64)).all()
                  and (got["met"] == met.cpu().numpy()).all() and (got["unmet"] == unmet.cpu().numpy()).all()
                  and (got["cov"] == cov.cpu().numpy()).all())
            sizes = [len(e) for e in every]
            print(json.dumps({"ok": bool(ok) and in_range, "world": world, "records": n, "sites": int(len(got)), "rows_per_rank": sizes,
                              "row_bounds": [int(x) for x in rb[:-1]], "timing": be.timing()}), flush=True)
        be.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    files = [f for grp_ in a.input_path for f in grp_]
    contigs = cf.parse_contigs_arg(a.contigs)
    fd.call_freq_distributed(files, a.prob_cf, a.result_file, a.sort, a.bed, a.gzip, contigs=contigs, grp=grp,
                             backend=backend, device=device if device is not None else 0)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
