"""Generate the feature-file fixtures by RUNNING THE REFERENCE reader (test infrastructure).

Run in the build container only (imports /root/reference):  python oracle/make_golden_features.py

* tests/golden/features_small.tsv.gz   -- a feature file: lines written by the reference's own
  ``_features_to_str`` (extract_features.py:381-395) from seeded synthetic features, plus a few
  hand-written lines with number spellings Python's float()/int() accept (exponents, signs,
  17-digit doubles, IUPAC bases);
* tests/golden/features_small_parsed.npz -- what the reference's ``_read_features_file``
  (call_modifications.py:55-127) put on its queue for that file, converted like ``FloatTensor``
  does (Python floats -> float32), and the batch sizes it cut (f5_batch_size = 7 reads).
The manifest entry ``features`` records digests.  Also checks ``feature_io.features_to_str``
against the reference writer."""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference, GOLD  # noqa: E402
from deepsignal_plant_b200 import synthetic, feature_io  # noqa: E402

EDGE = [
    # exponents, explicit signs, long decimals, integers spelled as floats, IUPAC bases
    "chrE\t7\t+\t7\tread_e1\tt\tACGTNWSMKRYBV\t" + ",".join(["1e-05", "-2.5E-3", "+0.5", "0.1234567890123456789", "3", "-0.0", "1e0"] + ["0.25"] * 6)
    + "\t" + ",".join(["0.300000011920929"] * 13) + "\t" + ",".join(str(i + 3) for i in range(13))
    + "\t" + ";".join(",".join(["0.1", "-1e-3", "2.5e+00", "7"] * 4) for _ in range(13)) + "\t1",
    "chrE\t8\t-\t99\tread_e1\tt\tDHZACGTACGTAC\t" + ",".join(["0.0"] * 13) + "\t" + ",".join(["1.0"] * 13) + "\t"
    + ",".join(["10"] * 13) + "\t" + ";".join(",".join(["0.000001"] * 16) for _ in range(13)) + "\t0",
]


class ListQueue:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)

    def qsize(self):
        return 0


def main():
    ref_models, ref_cm, ref_freq = import_reference()
    import deepsignal_plant.extract_features as ref_ex
    n, T, S = 240, 13, 16
    feats = synthetic.make_features(n, T, S, seed=31)
    info = synthetic.make_sampleinfo(n, seed=31)
    rng = np.random.default_rng(31)
    labels = rng.integers(0, 2, n)
    lines = []
    for i in range(n):
        w = info[i].split("\t")
        kmer = "".join(feature_io.code2base_dna[int(c)] for c in feats["kmer"][i])
        tup = (w[0], int(w[1]), w[2], int(w[3]), w[4], w[5], kmer,
               feats["base_means"][i].astype(np.float64), feats["base_stds"][i].astype(np.float64),
               feats["base_signal_lens"][i].astype(np.int64), feats["signals"][i].astype(np.float64).tolist(), int(labels[i]))
        line = ref_ex._features_to_str(tup)
        mine = feature_io.features_to_str(info[i], feats["kmer"][i], feats["base_means"][i].astype(np.float64),
                                          feats["base_stds"][i].astype(np.float64), feats["base_signal_lens"][i],
                                          feats["signals"][i].astype(np.float64), labels[i])
        assert mine == line, "feature_io.features_to_str differs from the reference writer"
        lines.append(line)
    lines[120:120] = EDGE
    text = "\n".join(lines) + "\n"
    path = os.path.join(GOLD, "features_small.tsv.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(text.encode())
    q = ListQueue()
    ref_cm._read_features_file(path, q, 7)
    assert q.items[-1] == "kill"
    batches = q.items[:-1]
    cat = lambda j, dt: np.concatenate([np.asarray(b[j], dtype=dt) for b in batches], 0)
    sampleinfo = [s for b in batches for s in b[0]]
    np.savez_compressed(os.path.join(GOLD, "features_small_parsed.npz"),
                        kmer=cat(1, np.float32), base_means=cat(2, np.float32), base_stds=cat(3, np.float32),
                        base_signal_lens=cat(4, np.float32), signals=cat(5, np.float32), labels=cat(6, np.int32),
                        batch_sizes=np.asarray([len(b[0]) for b in batches], np.int64))
    mpath = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(mpath))
    manifest["features"] = dict(n=len(lines), seq_len=T, signal_len=S, feature_seed=31, f5_batch_size=7,
                                text_sha256=hashlib.sha256(text.encode()).hexdigest(),
                                sampleinfo_sha256=hashlib.sha256("\n".join(sampleinfo).encode()).hexdigest(),
                                reference_batches=len(batches))
    with open(mpath, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("features golden: %d lines, %d reference batches, %d bytes of text" % (len(lines), len(batches), len(text)))


if __name__ == "__main__":
    main()
