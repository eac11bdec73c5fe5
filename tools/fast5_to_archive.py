#!/usr/bin/env python
"""Decode tombo-resquiggled single-read fast5 files ONCE into the decoded-reads archive that
`deepsignal_plant_b200 extract` / `call_mods` take in place of a fast5 directory.

    python tools/fast5_to_archive.py -i fast5_dir -o reads.npz [--corrected_group RawGenomeCorrected_000]
                                     [--basecall_subgroup BaseCalled_template] [--recursively yes]

NEEDS h5py, which this build image does not have: tests/test_fast5_converter.py exercises it against a
stand-in module (and, in the build container, runs the reference's own accessors on the same stand-in); it
has not been run on real fast5 files here.  It reads exactly the fields the reference's three accessors read (deepsignal_plant/extract_features.py):
  * `_get_alignment_info_from_fast5` (:150-176) / `_get_alignment_attrs_of_each_strand` (:94-129):
    Analyses/<corrected_group>/<basecall_subgroup>/Alignment attrs mapped_strand, mapped_chrom, mapped_start;
    strand 't' for a template subgroup, 'c' otherwise; reads without an Alignment group are skipped (:166-173);
  * `_get_readid_from_fast5` (:132-147): attrs['read_id'] of the first group under Raw/Reads;
  * `_get_label_raw` (:44-91): Raw/Reads/<first>/Signal (int16 DAC samples) and the Events table of the
    corrected group -- start (+ attrs['read_start_rel_to_raw'], :80), length, base;
  * `_get_scaling_of_a_read` (:255-273): UniqueGlobalKey/channel_id attrs range / digitisation and offset.
A file that cannot be read is counted and skipped, as `_extract_features` does (:373-375)."""
import argparse
import fnmatch
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _text(v):
    return v.decode("utf-8") if isinstance(v, (bytes, np.bytes_)) else str(v)


def decode_fast5(path, corrected_group, basecall_subgroup):
    import h5py
    with h5py.File(path, mode="r") as f:
        strand_path = "/".join(["Analyses", corrected_group, basecall_subgroup])
        if strand_path + "/Alignment" not in f:
            return None
        attrs = f[strand_path + "/Alignment"].attrs
        first = list(f["Raw/Reads"].keys())[0]
        read = f["Raw/Reads/" + first]
        events = f[strand_path + "/Events"]
        rel = events.attrs["read_start_rel_to_raw"]
        # a file without channel info raises KeyError here, which makes it an unreadable file below -- what happens in
        # the reference too (_get_scaling_of_a_read only catches IOError; _extract_features counts the read as an error)
        ch = f["UniqueGlobalKey/channel_id"].attrs
        scaling, offset = np.float64(ch["range"]) / np.float64(ch["digitisation"]), np.float64(ch["offset"])
        return dict(readname=_text(read.attrs["read_id"]), strand="t" if strand_path.endswith("template") else "c",
                    alignstrand=_text(attrs["mapped_strand"]), chrom=_text(attrs["mapped_chrom"]),
                    chrom_start=int(attrs["mapped_start"]), raw=np.asarray(read["Signal"][()]),
                    scaling=scaling, offset=offset,
                    ev_start=np.asarray(events["start"], np.int64) + int(rel), ev_len=np.asarray(events["length"], np.int64),
                    ev_base="".join(_text(b) for b in events["base"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fast5_dir", "-i", required=True)
    ap.add_argument("--write_path", "-o", required=True)
    ap.add_argument("--recursively", "-r", default="yes")
    ap.add_argument("--corrected_group", default="RawGenomeCorrected_000")
    ap.add_argument("--basecall_subgroup", default="BaseCalled_template")
    a = ap.parse_args()
    from deepsignal_plant_b200 import extract_features as ef
    files = []
    if a.recursively.lower() in ("yes", "true", "t", "1"):
        for root, _, names in os.walk(os.path.abspath(a.fast5_dir)):
            files += [os.path.join(root, n) for n in fnmatch.filter(names, "*.fast5")]
    else:
        files = [os.path.join(a.fast5_dir, n) for n in os.listdir(a.fast5_dir) if n.endswith(".fast5")]
    reads, errors = [], 0
    for p in files:
        try:
            rd = decode_fast5(p, a.corrected_group, a.basecall_subgroup)
        except Exception:                      # noqa: BLE001 -- the reference counts and skips (:373-375)
            errors += 1
            continue
        if rd is not None:
            if rd["raw"].dtype != np.int16:
                raise ValueError("%s: Raw/Reads Signal is %s, expected int16 DAC samples" % (p, rd["raw"].dtype))
            reads.append(rd)
    ef.save_reads(a.write_path, reads)
    print("%d fast5 files: %d reads written to %s, %d unreadable" % (len(files), len(reads), a.write_path, errors))


if __name__ == "__main__":
    main()
