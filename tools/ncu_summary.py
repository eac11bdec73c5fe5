#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, no GPU needed): one row per launch with the counters DESIGN.md
quotes, and -- with --traffic-json -- the dram bytes per site of the dominant kernel (lstm_comb layer 1,
layer_kernel<8,256,LSTM>) stamped with the sha256 of csrc/kernels_tc.cu, which bench.py checks before quoting it.

    python tools/ncu_summary.py gpurun_out/prof_r02.ncu-rep profiles/r02_ncu_full_summary.csv --batch 37888 \
        --traffic-json profiles/roofline_traffic.json
"""
import argparse
import csv
import hashlib
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "gpc__cycles_elapsed.max.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out_csv")
    ap.add_argument("--batch", type=int, default=37888)
    ap.add_argument("--traffic-json", default=None)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    keep = [m for m in METRICS if m in col]
    with open(a.out_csv, "w", newline="") as f:
        w = csv.writer(f)
        f.write("# %s\n" % (a.note or "ncu --set full --clock-control none, one forward step, batch %d; one row per launch" % a.batch))
        w.writerow(["Kernel Name", "Grid Size", "Block Size"] + keep)
        w.writerow(["", "", ""] + [units[col[m]] for m in keep])
        for r in body:
            w.writerow([r[col["Kernel Name"]][:110], r[col["Grid Size"]], r[col["Block Size"]]] + [r[col[m]] for m in keep])
    if a.traffic_json:
        # launch order of one step: prep, branch_fused, comb l0 (<4,256>), comb l1 (<8,256>), comb l2, head
        comb = [r for r in body if "layer_kernel<8, 256" in r[col["Kernel Name"]].replace("(int)", "")]
        r = comb[0]
        to_b = lambda v, u: float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd = to_b(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_b(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        sha = hashlib.sha256(open(os.path.join(ROOT, "deepsignal_plant_b200", "csrc", "kernels_tc.cu"), "rb").read()).hexdigest()
        json.dump({"kernel": "layer_kernel<8,256,MODE_LSTM> (lstm_comb layer 1)",
                   "source": "%s (ncu --set full, batch %d: dram__bytes_read.sum %.2f MB + dram__bytes_write.sum %.2f MB per launch)"
                             % (os.path.relpath(a.out_csv, ROOT), a.batch, rd / 1e6, wr / 1e6),
                   "kernels_tc_sha256": sha, "dram_bytes_per_site": int(round((rd + wr) / a.batch)),
                   "algorithmic_bytes_per_site": 26624,
                   "note": "algorithmic = 13 x 512 fp16 read + 13 x 512 fp16 written per site; the forward and the reverse direction CTAs "
                           "each stream the input image (second read mostly from L2)"}, open(a.traffic_json, "w"), indent=1)
        print("traffic: %.1f B/site (algorithmic 26624)" % ((rd + wr) / a.batch))


if __name__ == "__main__":
    main()
