"""Text boundary of call_mods: native feature-file parser and output formatter (host code in
libdsp_b200, no GPU needed) against what the reference's reader / writer loop produce."""
import gzip
import hashlib
import os

import numpy as np
import pytest

import cases
from deepsignal_plant_b200 import _native, feature_io, synthetic
from deepsignal_plant_b200 import call_modifications as cm

GOLD = cases.GOLD
FEAT = cases.MANIFEST["features"]
KEYS = ("kmer", "base_means", "base_stds", "base_signal_lens", "signals")


def gold_arrays():
    return np.load(os.path.join(GOLD, "features_small_parsed.npz"))


@pytest.mark.parametrize("batch_sites", [4096, 50, 1])
def test_parser_matches_reference_reader(batch_sites):
    g = gold_arrays()
    rd = feature_io.FeatureFileReader(os.path.join(GOLD, "features_small.tsv.gz"), FEAT["seq_len"], FEAT["signal_len"],
                                      batch_sites=batch_sites, pinned=False, slots=2, nthreads=3)
    got = {k: [] for k in KEYS + ("labels",)}
    info = []
    for b in rd:
        assert 1 <= b.n <= batch_sites
        for k, t in zip(KEYS, b.arrays()):
            got[k].append(t.numpy().copy())
        got["labels"].append(b.labels.numpy().copy())
        info += b.sampleinfo()
    assert len(info) == FEAT["n"] == rd.sites_read
    assert hashlib.sha256("\n".join(info).encode()).hexdigest() == FEAT["sampleinfo_sha256"]
    for k in got:
        a = np.concatenate(got[k], 0)
        assert a.dtype == g[k].dtype and a.shape == g[k].shape
        assert a.tobytes() == g[k].tobytes(), k          # bit-identical to float32(float(text))


def test_reference_list_view_and_text_digest():
    text = gzip.open(os.path.join(GOLD, "features_small.tsv.gz"), "rb").read()
    assert hashlib.sha256(text).hexdigest() == FEAT["text_sha256"]
    b = next(iter(feature_io.FeatureFileReader(os.path.join(GOLD, "features_small.tsv.gz"), 13, 16, pinned=False)))
    info, kmers, means, stds, lens, sig, labels = b.as_reference_lists()
    first = text.split(b"\n")[0].decode().split("\t")
    assert info[0] == "\t".join(first[:6])
    assert kmers[0] == [cm.base2code_dna[c] for c in first[6]] and lens[0] == [int(x) for x in first[9].split(",")]
    assert labels[0] == int(first[11]) and len(sig[0]) == 13 and len(sig[0][0]) == 16


@pytest.mark.parametrize("bad", ["too\tfew\tcolumns\n",
                                 "c\t1\t+\t1\tr\tt\tACGTACGTACGTA\t" + ",".join(["0.1"] * 12) + "\t" + ",".join(["0.1"] * 13) + "\t"
                                 + ",".join(["3"] * 13) + "\t" + ";".join([",".join(["0"] * 16)] * 13) + "\t1\n",     # 12 means
                                 "c\t1\t+\t1\tr\tt\tACGTACGTACGTX\t" + ",".join(["0.1"] * 13) + "\t" + ",".join(["0.1"] * 13) + "\t"
                                 + ",".join(["3"] * 13) + "\t" + ";".join([",".join(["0"] * 16)] * 13) + "\t1\n"])    # base X
def test_malformed_lines_fail_loudly(tmp_path, bad):
    p = tmp_path / "bad.tsv"
    p.write_text(bad)
    with pytest.raises(_native.DspError, match="feature file"):
        list(feature_io.FeatureFileReader(str(p), 13, 16, pinned=False))


def test_byte_range_shards_cover_the_file_once(tmp_path):
    text = gzip.open(os.path.join(GOLD, "features_small.tsv.gz"), "rb").read()
    p = tmp_path / "f.tsv"
    p.write_bytes(text)
    size = len(text)
    for world in (1, 2, 3, 7):
        info = []
        for r in range(world):
            rng = (size * r // world, size * (r + 1) // world)
            for b in feature_io.FeatureFileReader(str(p), 13, 16, batch_sites=64, pinned=False, byte_range=rng):
                info += b.sampleinfo()
        assert hashlib.sha256("\n".join(info).encode()).hexdigest() == FEAT["sampleinfo_sha256"], world
    with pytest.raises(ValueError):
        feature_io.FeatureFileReader(os.path.join(GOLD, "features_small.tsv.gz"), 13, 16, byte_range=(0, 10))


def test_formatter_reproduces_reference_call_lines(tmp_path):
    # the reference's _call_mods output (tests/golden/callmods_77.tsv.gz) from its own probabilities
    e = cases.MANIFEST["callmods"]
    n = 700
    feats = synthetic.make_features(e["n"], 13, 16, seed=e["feature_seed"])
    info = synthetic.make_sampleinfo(e["n"], seed=e["feature_seed"])
    p = tmp_path / "feat.tsv"
    with open(p, "w") as f:
        for i in range(n):
            f.write(feature_io.features_to_str(info[i], feats["kmer"][i], feats["base_means"][i], feats["base_stds"][i],
                                               feats["base_signal_lens"][i], feats["signals"][i], 0) + "\n")
    probs = np.load(os.path.join(GOLD, "callmods_%d_probs.npz" % e["rng_seed"]))["probs"][:n]
    gold = cases.read_gz("callmods_%d.tsv.gz" % e["rng_seed"]).splitlines()[:n]
    out = b""
    k = 0
    for b in feature_io.FeatureFileReader(str(p), 13, 16, batch_sites=256, pinned=False):
        pr = probs[k:k + b.n]
        out += feature_io.format_calls(b, pr, pr.argmax(1).astype(np.int32), nthreads=3)
        k += b.n
    assert out.decode().splitlines() == gold


def test_float32_text_matches_numpy_str():
    # str(numpy.float32) for the rounded, renormalised probabilities (call_modifications.py:177-188)
    rng = np.random.default_rng(3)
    p0 = np.concatenate([rng.random(20000), 10.0 ** rng.uniform(-9, -2, 20000), [0.0, 1.0, 0.5, 1e-6, 5.6e-05, 0.9999995]]).astype(np.float32)
    probs = np.stack([p0, (1 - p0).astype(np.float32)], 1)
    n = len(p0)
    info = b"c\t1\t+\t1\tr\tt"
    b = feature_io.FeatureBatch()
    b.n, b.seq_len = n, 13
    b.info_text = np.frombuffer(info * n, dtype=np.uint8)
    b.info_off = np.arange(n + 1, dtype=np.int64) * len(info)
    b.kmer = np.tile(np.array([0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0], np.float32), (n, 1))     # ACGTACGTACGTA
    got = feature_io.format_calls(b, probs, np.zeros(n, np.int32)).decode().splitlines()
    p0n, p1n = cm.normalise_probs(probs)
    for i in range(n):
        w = got[i].split("\t")
        assert w[6] == str(p0n[i]) and w[7] == str(p1n[i]), (i, w, p0n[i], p1n[i])
        assert w[8] == "0" and w[9] == "ACGTA"          # kmer[c-2:c+3], c = 6


def test_reference_style_reader_cuts_the_same_batches():
    # _read_features_file (call_modifications.py:55-127): same queue items and batch boundaries
    g = gold_arrays()
    q = cm.SimpleQueue()
    cm._read_features_file(os.path.join(GOLD, "features_small.tsv.gz"), q, FEAT["f5_batch_size"])
    items = []
    while not q.empty():
        items.append(q.get())
    assert items[-1] == "kill" and len(items) - 1 == FEAT["reference_batches"]
    assert [len(b[0]) for b in items[:-1]] == g["batch_sizes"].tolist()
    cat = lambda j, dt: np.concatenate([np.asarray(b[j], dtype=dt) for b in items[:-1]], 0)
    for j, key, dt in ((1, "kmer", np.float32), (2, "base_means", np.float32), (3, "base_stds", np.float32),
                       (4, "base_signal_lens", np.float32), (5, "signals", np.float32), (6, "labels", np.int32)):
        assert cat(j, dt).tobytes() == g[key].tobytes(), key
    assert isinstance(items[0][1][0][0], int) and isinstance(items[0][4][0][0], int) and isinstance(items[0][2][0][0], float)


def test_writer_worker(tmp_path):
    q = cm.SimpleQueue()
    q.put(["a\t1", "b\t2"])
    q.put(["c\t3"])
    q.put("kill")
    out = tmp_path / "o.tsv"
    cm._write_predstr_to_file(str(out), q, True)
    assert gzip.open(str(out) + ".gz", "rt").read() == "a\t1\nb\t2\nc\t3\n"


def test_parser_matches_python_float_on_random_spellings(tmp_path):
    # every number goes float(text) -> float32 in the reference (call_modifications.py:90-95, FloatTensor); the
    # parser's fast paths (exact integer / power of ten, eight fraction digits at a time) must agree bit for bit
    rng = np.random.default_rng(11)
    T, S, n = 13, 16, 300

    def spell(v, how):
        if how == 0:
            return repr(round(float(v), int(rng.integers(0, 8))))
        if how == 1:
            return "%.*f" % (int(rng.integers(0, 8)), v)
        if how == 2:
            return "%d" % int(v * 100)
        if how == 3:
            return "%.*e" % (int(rng.integers(0, 12)), v)
        if how == 4:
            return "%.17g" % v
        if how == 5:
            return "+%s" % abs(round(float(v), 3))
        if how == 6:
            return "%07.3f" % abs(v)
        return ["-0.0", "0.", ".5", "-.25", "1e-7", "12345678.9", "0.12345678", "1234567.1234567", "00.5", "9999999.9999999"][int(rng.integers(0, 10))]

    lines, want_m, want_s = [], [], []
    for i in range(n):
        vals = rng.normal(0, 1, T * 2 + T * S) * 10.0 ** rng.integers(-3, 4, T * 2 + T * S)
        toks = [spell(v, int(rng.integers(0, 8))) for v in vals]
        means, stds, sig = toks[:T], toks[T:2 * T], toks[2 * T:]
        lines.append("\t".join(["chr1", str(i), "+", str(i), "read%d" % (i // 7), "t", "ACGTACGTACGTA", ",".join(means), ",".join(stds),
                                ",".join(["5"] * T), ";".join(",".join(sig[t * S:(t + 1) * S]) for t in range(T)), "1"]))
        want_m.append([np.float32(float(x)) for x in means + stds])
        want_s.append([np.float32(float(x)) for x in sig])
    p = tmp_path / "rand.tsv"
    p.write_text("\n".join(lines) + "\n")
    got_m, got_s = [], []
    for b in feature_io.FeatureFileReader(str(p), T, S, batch_sites=64, pinned=False, nthreads=3):
        got_m.append(np.concatenate([b.base_means.numpy(), b.base_stds.numpy()], 1))
        got_s.append(b.signals.numpy().reshape(b.n, -1).copy())          # the slot is recycled a few batches later
    got_m, got_s = np.concatenate(got_m), np.concatenate(got_s)
    assert np.array_equal(got_m.view(np.uint32), np.array(want_m, np.float32).view(np.uint32))
    assert np.array_equal(got_s.view(np.uint32), np.array(want_s, np.float32).view(np.uint32))
