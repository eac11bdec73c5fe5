"""Generate the `call_freq --contigs` fixtures by RUNNING THE REFERENCE (test infrastructure).

    python oracle/make_golden_freq_contigs.py        # build container only (imports /root/reference)

Runs the reference's ``call_mods_frequency_to_file`` (call_mods_freq.py:218-296) in its per-contig
mode (``--contigs``, ``:262-295``: split by contig, one aggregation per contig, results concatenated
in sorted temp-file-name order) on 20 000 seeded synthetic call records, and stores the output
bytes under tests/golden/ together with a manifest entry."""
from __future__ import annotations

import argparse
import gzip
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference, GOLD  # noqa: E402
from deepsignal_plant_b200 import synthetic  # noqa: E402

CASES = [("names", "chr3,chr10,chr1,chr7,chrNone", False, False, 0.5),
         ("names_sorted_bed", "chr3,chr10,chr1,chr7,chrNone", True, True, 0.0),
         ("fasta", None, True, False, 0.2)]
FASTA = ">chr9 some description\nACGT\n>chr2\nAC\n>chr11\nGG\n>chrAbsent\nTT\n"


def main():
    _, _, ref_freq = import_reference()
    lines = synthetic.make_callmods_records(20000, n_chrom=12, n_pos=400, seed=6)
    manifest = {"input": dict(n=len(lines), seed=6, n_chrom=12, n_pos=400,
                              sha256=hashlib.sha256("\n".join(lines).encode()).hexdigest()), "fasta_text": FASTA}
    with tempfile.TemporaryDirectory() as tmp:
        a, b = os.path.join(tmp, "a.tsv"), os.path.join(tmp, "b.tsv.gz")
        with open(a, "w") as f:
            f.write("\n".join(lines[:9000]) + "\n")
        with gzip.open(b, "wt") as f:
            f.write("\n".join(lines[9000:]) + "\n")
        fa = os.path.join(tmp, "genome.fa")
        open(fa, "w").write(FASTA)
        for name, contigs, is_sort, is_bed, cf in CASES:
            out = os.path.join(tmp, "out_%s.txt" % name)
            args = argparse.Namespace(input_path=[a, b], result_file=out, prob_cf=cf, file_uid=None, sort=is_sort,
                                      bed=is_bed, gzip=False, contigs=contigs if contigs else fa, nproc=2)
            ref_freq.call_mods_frequency_to_file(args)
            data = open(out).read()
            with gzip.GzipFile(os.path.join(GOLD, "freq_contigs_%s.txt.gz" % name), "wb", mtime=0) as f:
                f.write(data.encode())
            manifest[name] = dict(contigs=contigs, sort=is_sort, bed=is_bed, prob_cf=cf, n_lines=data.count("\n"))
            print(name, data.count("\n"), "lines")
    mpath = os.path.join(GOLD, "manifest.json")
    m = json.load(open(mpath))
    m["freq_contigs"] = manifest
    with open(mpath, "w") as f:
        json.dump(m, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
