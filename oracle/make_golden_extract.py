"""Generate the feature-extraction fixtures by RUNNING THE REFERENCE (test infrastructure).

Run in the build container only (imports /root/reference):  python oracle/make_golden_extract.py

The reference's ``_extract_features`` (extract_features.py:280-378) is called unmodified.  h5py is
absent in this image, so its three fast5 accessors (``_get_alignment_info_from_fast5`` :150-176,
``_get_label_raw`` :37-91, ``_get_scaling_of_a_read`` :255-273) are monkeypatched to serve synthetic
decoded reads (``deepsignal_plant_b200.synthetic.make_reads``) keyed by a fake path; statsmodels is
absent too, so ``robust.mad`` is bound to the restatement of its published definition
(``oracle/extract_oracle.py::mad``) -- that one function is therefore not pinned by these fixtures.
Python's global ``random`` is seeded so that the ordered subsamples (:247-249) are reproducible.

Per case ``tests/golden/extract_<case>.npz`` holds the reads (concatenated arrays), what the reference
returned (float64 means/stds, lens, rectangles, k-mers, the six sample-info columns), the lines its
``_features_to_str`` (:381-395) wrote, and the subsample offsets (recovered by running the oracle on the
same random stream and checking it reproduces the reference's output exactly).
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference, GOLD  # noqa: E402
from oracle import extract_oracle as eo  # noqa: E402
from deepsignal_plant_b200 import synthetic  # noqa: E402

CASES = {
    # name: (reads kwargs, motifs, mod_loc, kmer_len, signals_len, with chrom2len, random seed[, normalize_method])
    "cg_13_16": (dict(n_reads=10, seed=5, mean_bases=160, long_every=3), "CG", 0, 13, 16, True, 12345),
    "chgchh_17_20": (dict(n_reads=6, seed=6, mean_bases=120, long_every=2, no_scaling_every=3), "CHG,CHH", 0, 17, 20,
                     False, 777),
    "cg_13_16_zscore": (dict(n_reads=8, seed=8, mean_bases=170, long_every=3, no_scaling_every=4), "CG", 0, 13, 16, True, 99,
                        "zscore"),
}


def main():
    import_reference()
    import deepsignal_plant.extract_features as ref_ex
    from deepsignal_plant.utils.process_utils import get_motif_seqs
    ref_ex.robust.mad = eo.mad
    manifest_path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(manifest_path))
    manifest["extract"] = {}
    for name, case in CASES.items():
        rkw, motifs, mod_loc, K, S, with_len, rseed = case[:7]
        method = case[7] if len(case) > 7 else "mad"
        reads = synthetic.make_reads(**rkw)
        table = {"/fake/%s.fast5" % r["readname"]: r for r in reads}
        ref_ex._get_alignment_info_from_fast5 = lambda fp, cg, bs: (
            table[fp]["readname"], table[fp]["strand"], table[fp]["alignstrand"], table[fp]["chrom"], table[fp]["chrom_start"])
        ref_ex._get_label_raw = lambda fp, cg, bs: (
            table[fp]["raw"], list(zip([int(x) for x in table[fp]["ev_start"]], table[fp]["ev_len"].astype(int),
                                       list(table[fp]["ev_base"]))))
        ref_ex._get_scaling_of_a_read = lambda fp: (table[fp]["scaling"], table[fp]["offset"])
        chrom2len = {"chr%d" % c: 200000 for c in range(1, 4)} if with_len else None
        motif_seqs = get_motif_seqs(motifs)
        assert sorted(motif_seqs) == sorted(eo.get_motif_seqs(motifs))
        random.seed(rseed)
        feats, err = ref_ex._extract_features(list(table), "RawGenomeCorrected_000", "BaseCalled_template", method,
                                              motif_seqs, mod_loc, chrom2len, K, S, 1, None, (None, None, None))
        assert err == 0 and len(feats) > 0
        mine, drawn = eo.extract_features(reads, method, motif_seqs, mod_loc, chrom2len, K, S, 1,
                                          rng=random.Random(rseed))
        assert len(mine) == len(feats)
        for a, b in zip(feats, mine):           # the oracle reproduces the reference bit for bit
            assert a[:7] == b[:7] and a[9] == b[9] and a[11] == b[11]
            assert np.array_equal(np.asarray(a[7]), np.asarray(b[7])) and np.array_equal(np.asarray(a[8]), np.asarray(b[8]))
            assert np.array_equal(np.asarray(a[10]), np.asarray(b[10]))
        lines = [ref_ex._features_to_str(f) for f in feats]
        out = eo.pack_reads(reads)
        out.update(
            info=np.array(["\t".join([f[0], str(f[1]), f[2], str(f[3]), f[4], f[5]]) for f in feats]),
            kmer=np.array([f[6] for f in feats]),
            means=np.array([f[7] for f in feats], np.float64), stds=np.array([f[8] for f in feats], np.float64),
            lens=np.array([f[9] for f in feats], np.int64), rect=np.array([f[10] for f in feats], np.float64),
            drawn=eo.drawn_to_array(drawn, K, S), lines=np.array(lines),
            motifs=np.array(motifs), mod_loc=np.int64(mod_loc), kmer_len=np.int64(K), signals_len=np.int64(S),
            chrom_len=np.int64(200000 if with_len else -1), random_seed=np.int64(rseed),
            normalize_method=np.array(method))
        np.savez_compressed(os.path.join(GOLD, "extract_%s.npz" % name), **out)
        n_long = int((out["lens"] > S).sum())
        manifest["extract"][name] = dict(sites=len(feats), reads=len(reads), samples=int(out["raw"].shape[0]),
                                         bases_longer_than_rect=n_long, max_dwell=int(out["lens"].max()),
                                         normalize_method=method)
        print(name, manifest["extract"][name])
    json.dump(manifest, open(manifest_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
