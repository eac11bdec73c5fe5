"""ORACLE (test infrastructure, not product code): time the UNMODIFIED reference's
``calculate_mods_frequency`` (``call_mods_freq.py:29-74``) on a synthetic call_mods file.

    python oracle/run_ref_callfreq.py --records 1000000 [--prob_cf 0.5]

Writes the file (same synthetic stream as ``freq_dist.synth_records``: 5 chromosomes, coverage ~20, 6-decimal
probabilities), runs the reference function on it in this process (single Python process, like the
reference's default mode) and prints one JSON line with records/s."""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=1000000)
    ap.add_argument("--coverage", type=int, default=20)
    ap.add_argument("--prob_cf", type=float, default=0.5)
    a = ap.parse_args()
    import torch
    from deepsignal_plant_b200 import freq_dist as fd
    from deepsignal_plant_b200 import call_mods_freq as cf
    from oracle import ref_import
    ref_freq = ref_import.import_reference("call_mods_freq")
    key, p0, p1, lab = (t.numpy() for t in fd.synth_records(0, a.records, max(a.records // a.coverage, 1), torch.device("cpu")))
    tmp = tempfile.mkdtemp(prefix="dsp_reffreq_")
    path = os.path.join(tmp, "calls.tsv")
    with open(path, "w") as f:
        for k, x, y, l in zip(key.tolist(), p0.tolist(), p1.tolist(), lab.tolist()):
            pos = k & ((1 << cf.POS_BITS) - 1)
            f.write("chr%d\t%d\t%s\t%d\tread\tt\t%r\t%r\t%d\tAACGT\n" % (k >> cf.POS_BITS, pos, "+-"[pos & 1], pos, x, y, l))
    t0 = time.perf_counter()
    table = ref_freq.calculate_mods_frequency([path], a.prob_cf)
    dt = time.perf_counter() - t0
    os.remove(path)
    os.rmdir(tmp)
    print(json.dumps({"records": a.records, "sites": len(table), "seconds": dt, "records_per_s": a.records / dt,
                      "reference": ref_import.ref_root()}))


if __name__ == "__main__":
    main()
