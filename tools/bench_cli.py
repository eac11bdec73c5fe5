#!/usr/bin/env python
"""End-to-end `call_mods` command line on a large synthetic feature file: text -> native parser ->
pinned batches -> CUDA forward -> native formatter -> output file.  Prints one JSON line.

    python tools/bench_cli.py [--sites 1000000]
    python tools/bench_cli.py --archive [--reads 2000]     # decoded-reads archive -> extract + call in one pass
    python tools/bench_cli.py --binary --sites 100000000   # binary feature file (.dspf) -> calls file, FRESH process
    python tools/bench_cli.py --binary --host-only         # the same pipeline with the device stubbed out (host ceiling, no GPU)"""
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
    ap.add_argument("--binary", action="store_true", help="input is a binary feature file (feature_bin.py); timed in a fresh process")
    ap.add_argument("--host-only", action="store_true", help="with --binary: stub the device out, in-process (no GPU needed)")
    ap.add_argument("--dir", default=None, help="where the synthetic files go (default: a temporary directory)")
    ap.add_argument("--variants", default="", help="with --binary: extra timed runs, comma-separated reader_threads:format_threads:stream_depth")
    ap.add_argument("--block-sites", type=int, default=65536, help="with --binary: distinct synthetic sites per block (tests use fewer)")
    ap.add_argument("--out-dir", default=None, help="with --binary: directory of the calls file (default: next to the input)")
    ap.add_argument("--keep-free-gb", type=float, default=8.0, help="with --binary: shrink --sites so that this much disk stays free")
    a = ap.parse_args()
    if a.binary:
        return binary(a)
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


class _NoDevice:
    """Stands in for ModelBiLSTM in --host-only runs: takes the batches, returns at once (prob 0.5 / label 0)."""
    num_classes = 2

    def submit_host(self, kmer, means, stds, lens, signals, logits, probs, labels):
        probs.numpy()[:] = 0.5
        labels.numpy()[:] = 0
        return 0

    def wait_host(self, ticket):
        return None


def binary(a):
    """`call_mods -i features.dspf`: a.sites synthetic sites (one 65 536-site block of distinct sites, repeated) ->
    calls file.  The command runs in a fresh interpreter: start-up, imports, CUDA context and model load are inside the
    wall clock.  The file has just been written, so it is read from the page cache."""
    import shutil
    import subprocess
    from deepsignal_plant_b200 import feature_bin
    block_n = max(1, int(a.block_sites))
    feats = synthetic.make_features(block_n, 13, 16, seed=1)
    info = synthetic.make_sampleinfo(block_n, seed=1)
    off = np.zeros(block_n + 1, np.int64)
    np.cumsum([len(x) for x in info], out=off[1:])
    text = np.frombuffer("".join(info).encode(), np.uint8)
    tmp = tempfile.mkdtemp(dir=a.dir)
    try:
        path, ckpt, out = os.path.join(tmp, "features.dspf"), os.path.join(tmp, "m.ckpt"), os.path.join(tmp, "calls.tsv")
        if a.out_dir:
            out = os.path.join(tempfile.mkdtemp(dir=a.out_dir), "calls.tsv")
        per_site = 1040 + 4 + 8 + text.size / block_n + (0 if a.out_dir else 90)            # input (+ output) bytes
        room = shutil.disk_usage(tmp).free - a.keep_free_gb * 2 ** 30
        sites = int(min(a.sites, max(block_n, room / per_site)))
        reps = max(1, sites // block_n)
        variants = []
        t0 = time.perf_counter()
        with feature_bin.FeatureBinWriter(path, 13, 16) as w:
            w.write(feats["kmer"], feats["base_means"], feats["base_stds"], feats["base_signal_lens"], feats["signals"], 0, text, off)
        if reps > 1:                                       # the same block again and again, written by several threads
            from concurrent.futures import ThreadPoolExecutor
            with open(path, "rb") as f:
                block = f.read()[feature_bin.HEADER_BYTES:]
            fd = os.open(path, os.O_WRONLY)
            try:
                os.ftruncate(fd, feature_bin.HEADER_BYTES + reps * len(block))
                def put(i):
                    view, at = memoryview(block), feature_bin.HEADER_BYTES + i * len(block)
                    while len(view):
                        k = os.pwrite(fd, view, at)
                        view, at = view[k:], at + k
                with ThreadPoolExecutor(8) as ex:
                    list(ex.map(put, range(1, reps)))
            finally:
                os.close(fd)
        t_write = time.perf_counter() - t0
        n = reps * block_n
        torch.manual_seed(1234)
        torch.save(ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True).state_dict(), ckpt)
        argv = ["call_mods", "-i", path, "-m", ckpt, "-o", out, "--host_threads", str(a.nproc)]
        env = dict(os.environ, DSP_B200_PROFILE="1", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
        if a.host_only:
            from deepsignal_plant_b200 import call_modifications as cm
            cm.load_model = lambda args, device=0: _NoDevice()
            if not torch.cuda.is_available():
                torch.Tensor.pin_memory = lambda self, *args, **kw: self         # pageable slots: this run times host code only
            os.environ["DSP_B200_PROFILE"] = "1"
            t0 = time.perf_counter()
            cli.main(argv)
            dt = time.perf_counter() - t0
            lines, first_s, clocks = [], None, None
        else:
            # an untimed first process on one block: the interpreter, torch and the CUDA libraries are paged in once per box
            warm = os.path.join(tmp, "warm.dspf")
            with feature_bin.FeatureBinWriter(warm, 13, 16) as w:
                w.write(feats["kmer"], feats["base_means"], feats["base_stds"], feats["base_signal_lens"], feats["signals"], 0, text, off)
            t0 = time.perf_counter()
            r = subprocess.run([sys.executable, "-m", "deepsignal_plant_b200", "call_mods", "-i", warm, "-m", ckpt, "-o", out + ".warm"],
                               env=env, capture_output=True, text=True)
            first_s = time.perf_counter() - t0
            if r.returncode != 0:
                raise SystemExit("call_mods failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
            import bench                                   # the repo's nvidia-smi sampler (clocks, power, throttle reasons)
            sampler = bench.ClockSampler(0)
            sampler.start()
            t0 = time.perf_counter()
            r = subprocess.run([sys.executable, "-m", "deepsignal_plant_b200"] + argv, env=env, capture_output=True, text=True)
            dt = time.perf_counter() - t0
            clocks = sampler.summary()
            clocks["power_w_series"] = [round(float(x[2])) for x in sampler.rows]
            if r.returncode != 0:
                raise SystemExit("call_mods failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
            lines = [l for l in r.stdout.splitlines() if "seconds" in l]
            for v in [x for x in a.variants.split(",") if x]:
                rt_, ft_, dp_ = v.split(":")
                t1 = time.perf_counter()
                rv = subprocess.run([sys.executable, "-m", "deepsignal_plant_b200"] + argv + ["--reader_threads", rt_, "--format_threads", ft_,
                                    "--stream_depth", dp_], env=env, capture_output=True, text=True)
                variants.append({"reader_threads:format_threads:stream_depth": v, "seconds": time.perf_counter() - t1, "rc": rv.returncode,
                                 "host_breakdown": [l for l in rv.stdout.splitlines() if "seconds" in l]})
        nout = 0
        with open(out, "rb") as f:
            while True:
                blk = f.read(1 << 26)
                if not blk:
                    break
                nout += blk.count(b"\n")
        print(json.dumps({"metric": "call_mods command line, binary feature file (.dspf) -> calls file (sites/s, wall clock of a fresh "
                                    "process incl. interpreter start, imports, CUDA context, model load)" if not a.host_only else
                                    "call_mods host pipeline with the device stubbed out (in-process)",
                          "sites": n, "sites_requested": a.sites, "lines_written": nout, "seconds": dt, "value": n / dt, "unit": "sites/s",
                          "input_bytes": os.path.getsize(path), "output_bytes": os.path.getsize(out), "host_threads": a.nproc,
                          "file_written_in_s": t_write, "one_block_process_s": first_s, "host_breakdown": lines, "clocks": clocks, "variants": variants}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if a.out_dir:
            shutil.rmtree(os.path.dirname(out), ignore_errors=True)


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
