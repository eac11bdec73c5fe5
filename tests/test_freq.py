"""call_freq: host logic on CPU; bit-exact GPU aggregation (-m gpu).  The multi-rank path is in test_freq_dist.py."""
import gzip
import os

import numpy as np
import pytest
import torch

import cases
from deepsignal_plant_b200 import call_mods_freq as cf
from deepsignal_plant_b200 import _native as _native_mod
from deepsignal_plant_b200 import synthetic
from oracle import freq_oracle

FREQ_CASES = sorted(k for k in cases.MANIFEST["freq"] if k.startswith("freq_"))


def inputs(key):
    if key == "edge":
        return open(cases.GOLD + "/freq_edge_input.tsv").read().splitlines()
    return synthetic.make_callmods_records(100000, n_chrom=12, n_pos=900, seed=5)


class _Row:
    def __init__(self, r):
        self._strand, self._pos_in_strand, self._kmer, self._prob_0, self._prob_1, self._met, self._unmet, self._coverage = r


def test_parse_matches_reference_field_rules(tmp_path):
    lines = inputs("edge")
    rec = cf.parse_lines(lines)
    assert len(rec) == len(lines) and rec.chrom[1] == "chr10" and rec.pos[1] == 5 and rec.label[1] == 1
    assert rec.p0[6] == float("5.6e-05") and rec.p1[7] == float("0.999999")
    p = tmp_path / "a.tsv.gz"
    with gzip.open(p, "wt") as f:
        f.write("\n".join(lines) + "\n")
    rec2 = cf.read_mods_file(str(p))
    for fld in ("chrom", "pos", "strand", "pos_in_strand", "p0", "p1", "label", "kmer"):
        assert (getattr(rec, fld) == getattr(rec2, fld)).all(), fld


@pytest.mark.parametrize("name", FREQ_CASES)
def test_render_from_oracle_table_matches_reference_bytes(name):
    # the writer half (write_sitekey2stats) on a table built by the oracle
    e = cases.MANIFEST["freq"][name]
    table = freq_oracle.aggregate(inputs(e["input"]), e["prob_cf"])
    mapping = {cf.key_sep.join([c, str(p)]): _Row(r) for (c, p), r in table.items()}
    assert cf.render_table(mapping, e["sort"], e["bed"]) == cases.read_gz(name + ".txt.gz")


def test_key_order_equals_python_tuple_order():
    chrom = np.array(["chr10", "chr2", "chrX", "chr1", "chr10"], dtype=object)
    pos = np.array([5, 100, 1, 7, 4], dtype=np.int64)
    ids, names = cf._chrom_ids(chrom)
    keys = cf.make_keys(ids, pos)
    want = sorted(range(5), key=lambda i: (chrom[i], int(pos[i])))
    assert np.argsort(keys, kind="stable").tolist() == want
    with pytest.raises(ValueError):
        cf.make_keys(ids, np.array([-1, 0, 0, 0, 0]))


def test_owner_of_key_is_balanced_and_deterministic():
    keys = cf.make_keys(np.repeat(np.arange(5), 1000), np.tile(np.arange(1000), 5))
    own = cf.owner_of_key(keys, 8)
    assert (own == cf.owner_of_key(keys, 8)).all()
    cnt = np.bincount(own, minlength=8)
    assert cnt.min() > 0.7 * len(keys) / 8 and cnt.max() < 1.3 * len(keys) / 8


# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", FREQ_CASES)
def test_gpu_freq_bit_exact_against_reference_bytes(name, tmp_path):
    e = cases.MANIFEST["freq"][name]
    lines = inputs(e["input"])
    half = len(lines) // 2
    a, b = tmp_path / "a.tsv", tmp_path / "b.tsv.gz"
    a.write_text("\n".join(lines[:half]) + "\n")
    with gzip.open(b, "wt") as f:
        f.write("\n".join(lines[half:]) + "\n")
    table = cf.calculate_mods_frequency([str(a), str(b)], e["prob_cf"])
    out = tmp_path / "out.txt"
    cf.write_sitekey2stats(table, str(out), e["sort"], e["bed"], False)
    assert out.read_text() == cases.read_gz(name + ".txt.gz")


@pytest.mark.gpu
def test_gpu_freq_dict_view_and_contig_filter():
    lines = inputs("edge")
    rec = cf.parse_lines(lines)
    t = cf.aggregate_records(rec, 0.0)
    want = freq_oracle.aggregate(lines, 0.0)
    assert t.keys() == [cf.key_sep.join([c, str(p)]) for c, p in want]
    row = t["chr2||100"]
    assert (row._strand, row._coverage, row._met, row._unmet) == ("+", 2, 1, 1)
    t2 = cf.aggregate_records(rec, 0.0, contig_name="chr1")
    assert cf.render_table(t2) == freq_oracle.render(freq_oracle.aggregate(lines, 0.0, "chr1"))
    assert len(cf.aggregate_records(cf.parse_lines([]), 0.0)) == 0


@pytest.mark.gpu
def test_gpu_freq_large_random_property():
    # size-independent properties at a size the Python oracle would need minutes for:
    # coverage sums to the callable count; met+unmet == coverage; permuting records of
    # DIFFERENT sites does not change any site's sums (order only matters within a site).
    rng = np.random.default_rng(1)
    n = 3_000_000
    pos = rng.integers(0, 200_000, n)
    chrom = rng.integers(0, 4, n)
    keys = cf.make_keys(chrom, pos)
    p1 = rng.random(n)
    p0 = 1.0 - p1
    label = (p1 > 0.5).astype(np.int32)
    k, first, s0, s1, met, unmet, cov = cf._aggregate_device(keys, p0, p1, label, 0.2, True, 0)
    callable_ = ~(np.abs(p0 - p1) < 0.2)
    assert cov.sum() == callable_.sum() and ((met + unmet) == cov).all()
    assert (np.diff(k.view(np.uint64).astype(np.int64)) > 0).all()
    # reference sums for a few sites, ordered replay on the host
    for j in rng.integers(0, len(k), 50):
        idx = np.nonzero((keys == k.view(np.uint64)[j]) & callable_)[0]
        a = 0.0
        for i in idx:
            a += p0[i]
        assert a == s0[j] and first[j] == idx[0] and cov[j] == len(idx)
    # stable partition by chromosome keeps within-site order -> identical results
    order = np.argsort(chrom, kind="stable")
    k2, _, s0b, s1b, *_ = cf._aggregate_device(keys[order], p0[order], p1[order], label[order], 0.2, True, 0)
    assert (k2 == k).all() and (s0b == s0).all() and (s1b == s1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["names", "names_sorted_bed", "fasta"])
def test_gpu_call_freq_contigs_mode_matches_reference_bytes(name, tmp_path):
    # `call_freq --contigs` (call_mods_freq.py:262-295) through the command line
    from deepsignal_plant_b200 import cli
    m = cases.MANIFEST["freq_contigs"]
    e = m[name]
    lines = synthetic.make_callmods_records(m["input"]["n"], n_chrom=m["input"]["n_chrom"], n_pos=m["input"]["n_pos"], seed=m["input"]["seed"])
    a, b = tmp_path / "a.tsv", tmp_path / "b.tsv.gz"
    a.write_text("\n".join(lines[:9000]) + "\n")
    with gzip.open(b, "wt") as f:
        f.write("\n".join(lines[9000:]) + "\n")
    contigs = e["contigs"]
    if contigs is None:
        fa = tmp_path / "genome.fa"
        fa.write_text(m["fasta_text"])
        contigs = str(fa)
    out = tmp_path / "freq.txt"
    argv = ["call_freq", "-i", str(a), "-i", str(b), "-o", str(out), "--prob_cf", str(e["prob_cf"]), "--contigs", contigs, "--nproc", "2"]
    argv += (["--sort"] if e["sort"] else []) + (["--bed"] if e["bed"] else [])
    assert cli.main(argv) == 0
    assert out.read_text() == cases.read_gz("freq_contigs_%s.txt.gz" % name)


def test_contig_argument_forms(tmp_path):
    assert cf.parse_contigs_arg(None) is None
    assert cf.parse_contigs_arg("chr3,chr10,chr1,chr3") == ["chr1", "chr10", "chr3"]
    names = tmp_path / "names.txt"
    names.write_text("chrB\nchrA\nchrB\n")
    assert cf.parse_contigs_arg(str(names)) == ["chrA", "chrB"]
    fa = tmp_path / "g.fasta"
    fa.write_text(">chr9 desc\nAC\n>chr2\nGT\n")
    assert cf.parse_contigs_arg(str(fa)) == ["chr9", "chr2"]


@pytest.mark.gpu
def test_gpu_freq_tables_combine_like_the_reference_script(tmp_path):
    # scripts/combine_call_mods_freq_files.py:25-42 merges per-file tables keyed by (chrom, pos, strand): counts add,
    # prob sums add (of %.3f-rounded values), rmet is recomputed.  Aggregating all files at once must agree with it:
    # integer columns exactly, the sums within the rounding of the per-file text.
    lines = synthetic.make_callmods_records(30000, n_chrom=4, n_pos=700, seed=9)
    parts = [lines[:9000], lines[9000:21000], lines[21000:]]
    files = []
    for i, p in enumerate(parts):
        f = tmp_path / ("calls%d.tsv" % i)
        f.write_text("\n".join(p) + "\n")
        files.append(str(f))
    combined = {}
    for f in files:
        out = tmp_path / "part.freq.txt"
        cf.write_sitekey2stats(cf.calculate_mods_frequency([f], 0.0), str(out), False, False, False)
        for line in out.read_text().splitlines():
            w = line.split("\t")
            key = (w[0], int(w[1]), w[2])
            c = combined.setdefault(key, [-1, 0.0, 0.0, 0, 0, 0, 0.0, "-"])
            c[0] = int(w[3]); c[1] += float(w[4]); c[2] += float(w[5]); c[3] += int(w[6]); c[4] += int(w[7]); c[5] += int(w[8])
            c[6] = c[3] / float(c[5]); c[7] = w[10]
    whole = tmp_path / "whole.freq.txt"
    cf.write_sitekey2stats(cf.calculate_mods_frequency(files, 0.0), str(whole), False, False, False)
    rows = [l.split("\t") for l in whole.read_text().splitlines()]
    assert len(rows) == len(combined) > 2000
    for w in rows:
        c = combined[(w[0], int(w[1]), w[2])]
        assert (int(w[6]), int(w[7]), int(w[8])) == (c[3], c[4], c[5])
        assert abs(float(w[4]) - c[1]) <= 0.0016 and abs(float(w[5]) - c[2]) <= 0.0016
        assert w[9] == "%.4f" % c[6]


def test_native_calls_parser_matches_python_field_rules(tmp_path):
    # dsp_parse_calls == ModRecord's field rules (parse_lines) on synthetic records, the edge-case file, extra
    # columns, CRLF / padded lines; compact Records behave like the object-array ones
    lines = synthetic.make_callmods_records(60000, n_chrom=12, n_pos=900, seed=3) + inputs("edge")
    lines[7] = lines[7] + "\textra\tcolumns"
    lines[8] = "  " + lines[8] + " \r"
    want = cf.parse_lines(lines)
    p = tmp_path / "calls.tsv"
    p.write_text("\n".join(lines) + "\n")
    got = cf._read_mods_file_native(str(p), nthreads=4)
    assert got is not None and got._codes is not None and len(got) == len(want)
    for fld in cf.Records.FIELDS:
        a, b = getattr(got, fld), getattr(want, fld)
        assert a.dtype == b.dtype and (a == b).all(), fld
    assert np.array_equal(got.p0.view(np.uint64), want.p0.view(np.uint64)) and np.array_equal(got.p1.view(np.uint64), want.p1.view(np.uint64))
    ids_a, names_a = got.chrom_ranks()
    ids_b, names_b = cf._chrom_ids(want.chrom)
    assert names_a == names_b and np.array_equal(ids_a, ids_b)
    wanted = {"chr10", "chr3", "nope"}
    keep = got.chrom_in(wanted)
    assert np.array_equal(keep, want.chrom_in(wanted)) and 0 < keep.sum() < len(got)
    sub = got.select(keep)
    assert (sub.chrom == want.chrom[keep]).all() and (sub.kmer == want.kmer[keep]).all() and sub._codes is not None
    idx = np.array([5, 0, 59999, 7])
    s_, q_, k_ = got.meta_at(idx)
    assert (s_ == want.strand[idx]).all() and (q_ == want.pos_in_strand[idx]).all() and (k_ == want.kmer[idx]).all()
    # two files with different chromosome tables concatenate into one table
    a, b = tmp_path / "a.tsv", tmp_path / "b.tsv.gz"
    a.write_text("\n".join(lines[:100]) + "\n")
    with gzip.open(b, "wt") as f:
        f.write("\n".join(lines[50000:50200]))                     # no trailing newline
    both = cf.Records.concat([cf.read_mods_file(str(a)), cf.read_mods_file(str(b))])
    ref = cf.parse_lines(lines[:100] + lines[50000:50200])
    assert both._codes is not None and all((getattr(both, f) == getattr(ref, f)).all() for f in cf.Records.FIELDS)
    # a k-mer column wider than the fixed cells takes the general path; malformed lines fail loudly
    wide = tmp_path / "wide.tsv"
    wide.write_text(lines[0].rsplit("\t", 1)[0] + "\t" + "ACGT" * 10 + "\n" + lines[1] + "\n")
    assert cf._read_mods_file_native(str(wide)) is None
    rec = cf.read_mods_file(str(wide))
    assert rec.kmer[0] == "ACGT" * 10 and rec.pos[1] == want.pos[1]
    for bad in ("chr1\t5\t+\n", lines[0].replace("\t0.", "\tx.", 1) + "\n", lines[0] + "\n\n" + lines[1] + "\n"):
        q = tmp_path / "bad.tsv"
        q.write_text(bad)
        with pytest.raises(_native_mod.DspError):
            cf._read_mods_file_native(str(q))
    e = tmp_path / "empty.tsv"
    e.write_text("")
    assert len(cf.read_mods_file(str(e))) == 0


def test_native_table_writer_equals_the_python_loop(monkeypatch):
    # dsp_format_freq vs the per-site Python formatting of write_sitekey2stats, on random rows (ties of %.3f / %.4f,
    # zero coverage, large sums) and both output formats; odd text columns take the Python loop
    rng = np.random.default_rng(5)
    n = 20000
    obj = lambda xs: np.asarray(xs, dtype=object)
    s0 = np.concatenate([rng.random(n - 6) * rng.choice([1, 50, 1e4], n - 6), [0.0005, 0.0015, 2.0005, 1e15, 0.00049999999999999994, 123456.7895]])
    cov = rng.integers(0, 60, n).astype(np.int32)
    met = (rng.random(n) * cov).astype(np.int32)
    t = cf.FreqTable(obj(["chr%d" % (i % 7) for i in range(n)]), rng.integers(0, 10 ** 9, n), obj(["+-"[i % 2] for i in range(n)]),
                     rng.integers(-1, 10 ** 9, n), obj(["ACGTN"[i % 5] * 5 for i in range(n)]), s0, s0[::-1].copy(), met,
                     (cov - met).astype(np.int32), cov, np.arange(n))
    for bed in (False, True):
        native = cf._render_native(t, bed)
        assert native is not None and native == cf.render_table(t, False, bed)
        with monkeypatch.context() as m:
            m.setattr(cf, "_render_native", lambda *a: None)
            assert cf.render_table(t, False, bed) == native
    assert cf._render_native(t.reorder(np.arange(0)), False) == ""
    odd = cf.FreqTable(obj(["chr\n1"]), np.array([5]), obj(["+"]), np.array([5]), obj(["AACGT"]), np.array([1.0]), np.array([0.0]),
                       np.array([1], np.int32), np.array([0], np.int32), np.array([1], np.int32), np.arange(1))
    assert cf._render_native(odd, False) is None and cf.render_table(odd).startswith("chr\n1\t5\t+")


@pytest.mark.gpu
@pytest.mark.parametrize("sort_by_key", [False, True])
def test_gpu_freq_key_hash_passes_equal_one_pass(sort_by_key):
    # more records than one dsp_freq_aggregate call takes (< 2^31) are aggregated in several passes over key-hash
    # shards (call_mods_freq._aggregate_device): same rows, same order, same float64 bits as one pass
    lines = synthetic.make_callmods_records(60000, n_chrom=7, n_pos=400, seed=21)
    rec = cf.parse_lines(lines)
    ids, names = cf._chrom_ids(rec.chrom)
    keys = cf.make_keys(ids, rec.pos)
    one = cf._aggregate_device(keys, rec.p0, rec.p1, rec.label, 0.2, sort_by_key, 0)
    many = cf._aggregate_device(keys, rec.p0, rec.p1, rec.label, 0.2, sort_by_key, 0, max_records=7000)
    assert len(one[0]) == len(many[0]) > 2000
    for a, b in zip(one, many):
        assert a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


# ---- bounded host memory: key-hash shards, the files re-read once per shard ------------------------------------------
def _standin_aggregate_records(rec, prob_cf, contig_name=None, sort_by_key=False, device=0):
    """TEST stand-in for the device aggregation (cf.aggregate_records) with the oracle's semantics -- float64 left to
    right per key in record order, rows in first-callable-appearance order -- so that the streaming host logic runs on a
    CPU box.  Lives in tests/: the package has no CPU path."""
    assert contig_name is None and not sort_by_key
    chrom, pos, p0, p1, lab = rec.chrom.tolist(), rec.pos.tolist(), rec.p0.tolist(), rec.p1.tolist(), rec.label.tolist()
    table = {}
    for i in range(len(rec)):
        if abs(p0[i] - p1[i]) < prob_cf:
            continue
        r = table.get((chrom[i], pos[i]))
        if r is None:
            r = table[(chrom[i], pos[i])] = [i, 0.0, 0.0, 0, 0]
        r[1] += p0[i]; r[2] += p1[i]
        r[3 if lab[i] == 1 else 4] += 1
    first = np.array([r[0] for r in table.values()], np.int64)
    strand, pis, kmer = rec.meta_at(first)
    col = lambda j, dt: np.array([r[j] for r in table.values()], dt)
    return cf.FreqTable(np.array([k[0] for k in table], object), np.array([k[1] for k in table], np.int64), strand, pis, kmer,
                        col(1, np.float64), col(2, np.float64), col(3, np.int32), col(4, np.int32),
                        (col(3, np.int32) + col(4, np.int32)), first, len(rec), int(sum(r[3] + r[4] for r in table.values())))


def _two_files(tmp_path, lines):
    third = len(lines) // 3
    a, b, c = tmp_path / "a.tsv", tmp_path / "b.tsv.gz", tmp_path / "c.tsv"
    a.write_text("\n".join(lines[:third]) + "\n")
    with gzip.open(b, "wt") as f:
        f.write("\n".join(lines[third:2 * third]) + "\n")
    c.write_text("\n".join(lines[2 * third:]) + "\n")
    return [str(a), str(b), str(c)]


@pytest.mark.parametrize("name", FREQ_CASES)
@pytest.mark.parametrize("shards,chunk", [(1, 1 << 30), (3, 700_000), (8, 40_000)])
def test_streaming_shards_give_the_reference_bytes_with_a_standin_device(name, shards, chunk, tmp_path, monkeypatch):
    monkeypatch.setattr(cf, "aggregate_records", _standin_aggregate_records)
    e = cases.MANIFEST["freq"][name]
    lines = inputs(e["input"])
    files = _two_files(tmp_path, lines)
    t = cf.calculate_mods_frequency_streaming(files, e["prob_cf"], shards=shards, chunk_bytes=chunk)
    assert t.n_records == len(lines)
    assert cf.render_table(t, e["sort"], e["bed"]) == cases.read_gz(name + ".txt.gz")


def test_streaming_is_chosen_by_the_record_budget_and_filters_contigs(tmp_path, monkeypatch, capsys):
    monkeypatch.setattr(cf, "aggregate_records", _standin_aggregate_records)
    monkeypatch.setattr(cf, "_warm_device_in_background", lambda device: None)
    lines = inputs("synth")
    files = _two_files(tmp_path, lines)
    est = cf.estimate_records(files)
    assert 0.8 * len(lines) < est < 2.5 * len(lines)          # the .gz member counts 5 x its size
    t = cf.calculate_mods_frequency(files, 0.5, max_host_records=20000)           # ~100 000 records: several shards
    assert "calls used.." in capsys.readouterr().out
    assert cf.render_table(t) == freq_oracle.render(freq_oracle.aggregate(lines, 0.5))
    one = cf.calculate_mods_frequency(files, 0.5, contig_name="chr3", max_host_records=20000)
    assert cf.render_table(one) == freq_oracle.render(freq_oracle.aggregate(lines, 0.5, "chr3")) and one.n_records == len(lines)
    assert len(cf.calculate_mods_frequency_streaming(files, 0.5, contigs={"nothing"}, shards=2)) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["freq_synth_cf0p5_unsorted_tsv", "freq_edge_cf0p0_sorted_bed"])
def test_gpu_streaming_call_freq_equals_reference_bytes(name, tmp_path):
    # the same through the real device aggregation and the command line: --max_host_records far below the input size
    from deepsignal_plant_b200 import cli
    e = cases.MANIFEST["freq"][name]
    lines = inputs(e["input"])
    files = _two_files(tmp_path, lines)
    t = cf.calculate_mods_frequency_streaming(files, e["prob_cf"], shards=5, chunk_bytes=300_000)
    assert cf.render_table(t, e["sort"], e["bed"]) == cases.read_gz(name + ".txt.gz")
    out = tmp_path / "out.txt"
    argv = ["call_freq", "-o", str(out), "--prob_cf", str(e["prob_cf"]), "--max_host_records", str(max(4, len(lines) // 7))]
    for f in files:
        argv += ["-i", f]
    assert cli.main(argv + (["--sort"] if e["sort"] else []) + (["--bed"] if e["bed"] else [])) == 0
    assert out.read_text() == cases.read_gz(name + ".txt.gz")


@pytest.mark.gpu
def test_gpu_streaming_contigs_mode_equals_the_in_memory_mode(tmp_path):
    from deepsignal_plant_b200 import cli
    lines = inputs("synth")
    files = _two_files(tmp_path, lines)
    outs = []
    for extra in ([], ["--max_host_records", "9000"]):
        out = tmp_path / ("o%d.txt" % len(outs))
        argv = ["call_freq", "-o", str(out), "--prob_cf", "0.2", "--contigs", "chr3,chr11,chr1", "--sort"] + extra
        for f in files:
            argv += ["-i", f]
        assert cli.main(argv) == 0
        outs.append(out.read_text())
    assert outs[0] == outs[1] and outs[0].count("\n") > 100
