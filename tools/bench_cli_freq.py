#!/usr/bin/env python
"""Wall clock of the real `call_freq` command line in a fresh process on a large calls file, with the host-side
breakdown (parse / aggregate / render + write), and the reference's command on a sample of the same file.

    python tools/bench_cli_freq.py [--records 50000000] [--ranks 1]

The file is a block of synthetic call_mods lines repeated to the requested size (page cache warm: the run measures
parsing, not the disk)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=50_000_000)
    ap.add_argument("--block", type=int, default=1_000_000)
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--ref_sample", type=int, default=1_000_000)
    a = ap.parse_args()
    from deepsignal_plant_b200 import synthetic
    tmp = tempfile.mkdtemp(prefix="dsp_clifreq_")
    block = "\n".join(synthetic.make_callmods_records(a.block, n_chrom=5, n_pos=a.block // 100, seed=3)) + "\n"
    path = os.path.join(tmp, "calls.tsv")
    reps = max(1, a.records // a.block)
    with open(path, "w") as f:
        for _ in range(reps):
            f.write(block)
    sample = os.path.join(tmp, "sample.tsv")
    if a.ref_sample > 0:                                        # the first ref_sample lines of the block (whole lines)
        lines = block.split("\n")[:min(a.ref_sample, a.block)]
        with open(sample, "w") as f:
            f.write("\n".join(lines) + "\n")
    n = reps * a.block
    out = os.path.join(tmp, "freq.tsv")
    tail = ["-m", "deepsignal_plant_b200", "call_freq", "-i", path, "-o", out, "--sort"]
    cmd = [sys.executable] + tail if a.ranks == 1 else \
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.ranks), "--master-addr", "127.0.0.1",
         "--master-port", "29741"] + tail
    env = dict(os.environ, DSP_B200_PROFILE="1")
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=env)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    sites = sum(1 for _ in open(out))
    res = {"command": "call_freq --sort, %d rank(s), fresh process" % a.ranks, "records": n, "file_GB": os.path.getsize(path) / 1e9,
           "sites": sites, "wall_s": dt, "records_per_s": n / dt, "host_breakdown": [l for l in r.stdout.splitlines() if "seconds" in l]}
    from oracle import ref_import
    if ref_import.available() and a.ref_sample > 0:
        code = ("import sys, time; sys.path.insert(0, %r)\nfrom oracle import ref_import\nm = ref_import.import_reference('call_mods_freq')\n"
                "t0 = time.perf_counter(); t = m.calculate_mods_frequency([%r], 0.5); m.write_sitekey2stats(t, %r, True, False, False)\n"
                "print('REF', time.perf_counter() - t0)\n" % (ROOT, sample, out + ".ref"))
        rr = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=ref_import.cpu_env())
        if rr.returncode == 0:
            sec = float([l for l in rr.stdout.splitlines() if l.startswith("REF")][-1].split()[1])
            m = sum(1 for _ in open(sample))
            res["reference"] = {"records": m, "wall_s": sec, "records_per_s": m / sec, "what": "calculate_mods_frequency + write_sitekey2stats, one process"}
    print(json.dumps(res), flush=True)
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
