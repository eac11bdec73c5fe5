"""Feature extraction from decoded reads: oracle and host bookkeeping vs the
fixtures the reference's own ``_extract_features`` produced (CPU); ``dsp_extract_features`` vs the same
fixtures and vs the oracle on fresh seeds, bit for bit (-m gpu)."""
import random

import numpy as np
import pytest
import torch

import cases
from deepsignal_plant_b200 import _native, synthetic
from deepsignal_plant_b200 import extract_features as ef
from oracle import extract_oracle as eo

EXTRACT_CASES = sorted(cases.MANIFEST["extract"])


def load(name):
    z = np.load(cases.GOLD + "/extract_%s.npz" % name)
    reads = eo.unpack_reads(z)
    K, S = int(z["kmer_len"]), int(z["signals_len"])
    chrom2len = None if int(z["chrom_len"]) < 0 else {"chr%d" % c: int(z["chrom_len"]) for c in range(1, 4)}
    return z, reads, K, S, chrom2len, eo.get_motif_seqs(str(z["motifs"])), int(z["mod_loc"])


def method_of(z):
    return str(z["normalize_method"])


# ---------------------------------------------------------------------------- CPU: oracle + host logic
@pytest.mark.parametrize("name", EXTRACT_CASES)
def test_oracle_reproduces_reference_fixture(name):
    z, reads, K, S, chrom2len, motif_seqs, mod_loc = load(name)
    feats, drawn = eo.extract_features(reads, method_of(z), motif_seqs, mod_loc, chrom2len, K, S, 1,
                                       rng=random.Random(int(z["random_seed"])))
    assert len(feats) == cases.MANIFEST["extract"][name]["sites"] == len(z["info"])
    assert ["\t".join([f[0], str(f[1]), f[2], str(f[3]), f[4], f[5]]) for f in feats] == list(z["info"])
    assert [f[6] for f in feats] == list(z["kmer"])
    assert np.array_equal(np.array([f[7] for f in feats]), z["means"])          # float64, bit for bit
    assert np.array_equal(np.array([f[8] for f in feats]), z["stds"])
    assert np.array_equal(np.array([f[9] for f in feats]), z["lens"])
    assert np.array_equal(np.array([f[10] for f in feats]), z["rect"])
    assert np.array_equal(eo.drawn_to_array(drawn, K, S), z["drawn"])
    assert (z["lens"] > S).sum() == cases.MANIFEST["extract"][name]["bases_longer_than_rect"] > 0
    assert z["lens"].max() > 128                                                # numpy's recursive pairwise form is covered


@pytest.mark.parametrize("name", EXTRACT_CASES)
def test_feature_lines_of_the_fixture_round_trip(name):
    # the lines the reference's _features_to_str wrote carry the 6-decimal means/stds: features_to_arrays(round_stats=True)
    from oracle import features_oracle
    z, reads, K, S, chrom2len, motif_seqs, mod_loc = load(name)
    feats, _ = eo.extract_features(reads, method_of(z), motif_seqs, mod_loc, chrom2len, K, S, 1,
                                   rng=random.Random(int(z["random_seed"])))
    arr = eo.features_to_arrays(feats, round_stats=True)
    info, kmers, means, stds, lens, sig, labels = features_oracle.read_features([str(x) for x in z["lines"]])
    assert info == list(z["info"])
    assert np.array_equal(np.asarray(kmers, np.float32), arr["kmer"])
    assert np.array_equal(np.asarray(means, np.float32), arr["base_means"])
    assert np.array_equal(np.asarray(stds, np.float32), arr["base_stds"])
    assert np.array_equal(np.asarray(lens, np.float32), arr["base_signal_lens"])
    assert np.array_equal(np.asarray(sig, np.float32), arr["signals"])


@pytest.mark.parametrize("name", EXTRACT_CASES)
def test_find_sites_and_sampleinfo_match_reference(name):
    z, reads, K, S, chrom2len, motif_seqs, mod_loc = load(name)
    batch = ef.pack_reads(reads)
    assert sorted(ef.get_motif_seqs(str(z["motifs"]))) == sorted(motif_seqs)
    sites = ef.find_sites(batch, motif_seqs, mod_loc, chrom2len, K)
    assert ef.sampleinfo(batch, sites) == list(z["info"])
    nb = (K - 1) // 2
    assert [bytes(batch.ev_base[e - nb:e + nb + 1]).decode() for e in sites.site_ev] == list(z["kmer"])


def test_find_sites_filters_match_oracle():
    reads = synthetic.make_reads(25, seed=11, mean_bases=90)
    chrom2len = {"chr1": 200000, "chr2": 200000}                      # chr3 missing -> pos_in_strand -1 (:331-335)
    motif_seqs = eo.get_motif_seqs("CHG")
    batch = ef.pack_reads(reads)
    for mod_loc, K in ((0, 13), (2, 9)):
        full, _ = eo.extract_features(reads, "mad", motif_seqs, mod_loc, chrom2len, K, 16, 1, rng=random.Random(1))
        want = ["\t".join([f[0], str(f[1]), f[2], str(f[3]), f[4], f[5]]) for f in full]
        assert ef.sampleinfo(batch, ef.find_sites(batch, motif_seqs, mod_loc, chrom2len, K)) == want and len(want) > 50
    full, _ = eo.extract_features(reads, "mad", motif_seqs, 0, chrom2len, 13, 16, 1, rng=random.Random(1))
    some = full[len(full) // 2]
    for region in (("chr2", None, None), (some[0], some[1] - 40, some[1] + 25), ("chrNone", None, None), (some[0], some[1], None)):
        sub, _ = eo.extract_features(reads, "mad", motif_seqs, 0, chrom2len, 13, 16, 1, regioninfo=region, rng=random.Random(1))
        got = ef.sampleinfo(batch, ef.find_sites(batch, motif_seqs, 0, chrom2len, 13, regioninfo=region))
        assert got == ["\t".join([f[0], str(f[1]), f[2], str(f[3]), f[4], f[5]]) for f in sub], region
    positions = {"||".join([f[0], str(f[1]), f[2]]) for f in full[::3]}
    sub, _ = eo.extract_features(reads, "mad", motif_seqs, 0, chrom2len, 13, 16, 1, positions=positions, rng=random.Random(1))
    got = ef.sampleinfo(batch, ef.find_sites(batch, motif_seqs, 0, chrom2len, 13, positions=positions))
    assert got == ["\t".join([f[0], str(f[1]), f[2], str(f[3]), f[4], f[5]]) for f in sub] and 0 < len(got) < len(full)


def test_sampleinfo_native_formatter_and_archive_round_trip(tmp_path):
    reads = synthetic.make_reads(9, seed=4, mean_bases=150)
    reads[2]["chrom_start"] = 0
    batch = ef.pack_reads(reads)
    for chrom2len in (None, {"chr1": 200000, "chr3": 123456789012}):
        sites = ef.find_sites(batch, ["CG"], 0, chrom2len, 13)
        text, off = ef.sampleinfo_packed(batch, sites, nthreads=3)
        want = ef.sampleinfo(batch, sites)
        assert [text[off[i]:off[i + 1]].tobytes().decode() for i in range(len(sites))] == want and len(want) > 20
    none = ef.find_sites(batch, ["ACGTACGTACGTTTTT"], 0, None, 13)
    text, off = ef.sampleinfo_packed(batch, none)
    assert text.size == 0 and off.tolist() == [0]
    path = str(tmp_path / "reads.npz")
    ef.save_reads(path, reads)
    again = ef.load_reads(path)
    for k, v in batch.arrays().items():
        assert np.array_equal(v, again.arrays()[k]), k
    part = again.slice(3, 7)
    sub = ef.find_sites(part, ["CG"], 0, None, 13)
    full = ef.find_sites(batch, ["CG"], 0, None, 13)
    keep = (full.site_read >= 3) & (full.site_read < 7)
    assert ef.sampleinfo(part, sub) == [x for x, k in zip(ef.sampleinfo(batch, full), keep) if k]
    np.savez(str(tmp_path / "bad.npz"), raw=np.zeros(3, np.int16))
    with pytest.raises(ValueError, match="not a decoded-reads archive"):
        ef.load_reads(str(tmp_path / "bad.npz"))
    assert ef.parse_region_str("chr1:5-9") == ("chr1", 5, 9) and ef.parse_region_str("chr1:5") == ("chr1", 5, None)
    assert ef.parse_region_str("chr1") == ("chr1", None, None) and ef.parse_region_str(None) == (None, None, None)
    with pytest.raises(ValueError, match="--region not set right"):
        ef.parse_region_str("chr1:a-b")


@pytest.mark.parametrize("name", EXTRACT_CASES)
def test_native_feature_line_formatter_reproduces_reference_lines(name):
    # dsp_format_features on the reference's own float64 values == the lines its _features_to_str wrote
    z, reads, K, S, chrom2len, motif_seqs, mod_loc = load(name)
    batch = ef.pack_reads(reads)
    sites = ef.find_sites(batch, motif_seqs, mod_loc, chrom2len, K)
    text = ef.format_features(batch, sites, np.around(z["means"], 6), np.around(z["stds"], 6), z["lens"].astype(np.float64),
                              z["rect"], 1, nthreads=3)
    assert text.decode().splitlines() == [str(x) for x in z["lines"]]
    # number spellings: str(numpy.float64) for the cases the fixtures do not contain
    vals = np.array([0.0, -0.0, 1.0, -1.5, 1e-4, 9.9e-05, 1e-06, 2.5e-05, 123456.789012, 1e15, 1e16, 1.5e16, 12345678.9,
                     0.1 + 0.2, 1 / 3, 5e-324, 1.7976931348623157e308, 100.0, 0.000123])
    one = ef.Sites(sites.site_read[:1], sites.site_ev[:1], sites.pos[:1], sites.pos_in_strand[:1])
    m = np.resize(vals, (1, K))
    sig = np.resize(vals, (1, K, S))
    line = ef.format_features(batch, one, m, m, np.full((1, K), 7.0), sig, 0).decode().rstrip("\n").split("\t")
    assert line[7] == ",".join(str(np.float64(v)) for v in m[0]) and line[9] == ",".join(["7"] * K) and line[11] == "0"
    assert line[10] == ";".join(",".join(str(np.float64(v)) for v in row) for row in sig[0])
    # randomised: 6-decimal values of every magnitude (the fast path) and raw doubles (shortest round-trip digits)
    rng = np.random.default_rng(3)
    n = 400
    many = ef.Sites(np.resize(sites.site_read, n), np.resize(sites.site_ev, n), np.resize(sites.pos, n), np.resize(sites.pos_in_strand, n))
    mag = 10.0 ** rng.integers(-7, 9, (n, K, S))
    sig = np.around(rng.normal(0, 1, (n, K, S)) * mag, 6)
    raw = rng.normal(0, 1, (n, K)) * 10.0 ** rng.integers(-12, 18, (n, K))
    lines = ef.format_features(batch, many, raw, np.around(raw, 6), np.full((n, K), 3.0), sig, 1, nthreads=4).decode().splitlines()
    assert len(lines) == n
    for i in (0, 1, 57, 199, 398, 399):
        w = lines[i].split("\t")
        assert w[7] == ",".join(str(v) for v in raw[i]) and w[8] == ",".join(str(v) for v in np.around(raw[i], 6))
        assert w[10] == ";".join(",".join(str(v) for v in row) for row in sig[i])
    assert "\n".join(lines) + "\n" == ef.format_features(batch, many, raw, np.around(raw, 6), np.full((n, K), 3.0), sig, 1, nthreads=1).decode()


def test_host_argument_errors():
    reads = synthetic.make_reads(3, seed=2, mean_bases=60)
    batch = ef.pack_reads(reads)
    with pytest.raises(ValueError, match="kmer_len must be odd"):
        ef.find_sites(batch, ["CG"], 0, None, 12)
    bad = dict(reads[0]); bad["ev_base"] = "x" + bad["ev_base"][1:]
    with pytest.raises(KeyError):
        ef.pack_reads([bad])
    bad = dict(reads[0]); bad["ev_len"] = bad["ev_len"].copy(); bad["ev_len"][-1] += 10 ** 6
    with pytest.raises(ValueError, match="outside"):
        ef.pack_reads([bad])
    sites = ef.find_sites(batch, ["CG"], 0, None, 13)
    with pytest.raises(ValueError):
        ef.extract_tensors(batch, sites, normalize_method="median")
    if not torch.cuda.is_available():
        with pytest.raises(_native.DspError, match="no CPU fallback"):
            ef.extract_tensors(batch, sites)


# ---------------------------------------------------------------------------- GPU: bit-exact parity
def _check_against(feats, drawn, batch, sites, K, S, round_stats, method="mad"):
    want = eo.features_to_arrays(feats, round_stats=round_stats)
    got = ef.extract_tensors(batch, sites, K, S, normalize_method=method, round_stats=round_stats,
                             drawn=eo.drawn_to_array(drawn, K, S))
    for k in ("kmer", "base_signal_lens", "base_means", "base_stds", "signals"):
        g = got[k].cpu().numpy()
        assert g.dtype == np.float32 and g.shape == want[k].shape, k
        assert np.array_equal(g.view(np.uint32), want[k].view(np.uint32)), (k, np.abs(g - want[k]).max())
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("name", EXTRACT_CASES)
@pytest.mark.parametrize("round_stats", [False, True])
def test_gpu_extract_matches_reference_fixture_bit_for_bit(name, round_stats):
    z, reads, K, S, chrom2len, motif_seqs, mod_loc = load(name)
    batch = ef.pack_reads(reads)
    sites = ef.find_sites(batch, motif_seqs, mod_loc, chrom2len, K)
    got = ef.extract_tensors(batch, sites, K, S, normalize_method=method_of(z), round_stats=round_stats, drawn=z["drawn"])
    means, stds = (np.around(z["means"], 6), np.around(z["stds"], 6)) if round_stats else (z["means"], z["stds"])
    eq = lambda a, b: np.array_equal(a.cpu().numpy().view(np.uint32), np.asarray(b, np.float32).view(np.uint32))
    assert eq(got["base_means"], means) and eq(got["base_stds"], stds)
    assert eq(got["base_signal_lens"], z["lens"]) and eq(got["signals"], z["rect"])
    assert eq(got["kmer"], [[eo.base2code_dna[c] for c in k] for k in z["kmer"]])
    # the per-read shift / scale _normalize_signals used, float64 bit for bit
    for i, rd in enumerate(reads):
        x = rd["raw"] if rd["scaling"] is None else eo.rescale_signals(rd["raw"], rd["scaling"], rd["offset"])
        want = (np.median(x), eo.mad(x)) if method_of(z) == "mad" else (np.mean(x), np.std(x))
        assert got["read_shift"][i].item() == float(want[0]) and got["read_scale"][i].item() == float(want[1])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,K,S,motifs", [(101, 13, 16, "CG"), (102, 17, 20, "CHH"), (103, 5, 8, "C"), (104, 13, 16, "CWG,CCG")])
def test_gpu_extract_matches_oracle_on_fresh_reads(seed, K, S, motifs):
    reads = synthetic.make_reads(60, seed=seed, mean_bases=220, long_every=4, no_scaling_every=5, stall_every=6)
    # edge reads: constant signal (MAD 0 -> samples kept as they are, :186-187), odd and even sample counts,
    # two-sample events everywhere, a read with a negative scaling (monotone decreasing rescale)
    flat = dict(reads[0]); flat["raw"] = np.full_like(flat["raw"], 431); flat["readname"] = "flat"
    odd = dict(reads[1]); odd["raw"] = odd["raw"][:len(odd["raw"]) - (len(odd["raw"]) % 2 == 0)]
    last = odd["ev_start"] + odd["ev_len"] <= len(odd["raw"])
    odd["ev_start"], odd["ev_len"], odd["ev_base"] = odd["ev_start"][last], odd["ev_len"][last], "".join(np.array(list(odd["ev_base"]))[last])
    odd["readname"] = "odd"
    neg = dict(reads[2]); neg["scaling"] = np.float64(-0.173); neg["readname"] = "neg"
    # DAC values spanning more levels than the per-read histogram holds: the select runs over the samples
    wide = dict(reads[3]); wide["readname"] = "wide"
    wide["raw"] = np.clip((wide["raw"].astype(np.int32) - int(np.median(wide["raw"]))) * 90, -32768, 32767).astype(np.int16)
    assert int(wide["raw"].max()) - int(wide["raw"].min()) > 20000
    reads = reads + [wide, flat, odd, neg]
    assert len(odd["raw"]) % 2 == 1
    motif_seqs = eo.get_motif_seqs(motifs)
    chrom2len = {"chr%d" % c: 200000 for c in range(1, 4)}
    batch = ef.pack_reads(reads)
    sites = ef.find_sites(batch, motif_seqs, 0, chrom2len, K)
    feats, drawn = eo.extract_features(reads, "mad", motif_seqs, 0, chrom2len, K, S, 1, rng=random.Random(seed))
    fz, dz = eo.extract_features(reads[:30] + reads[-4:], "zscore", motif_seqs, 0, chrom2len, K, S, 1, rng=random.Random(seed))
    bz = ef.pack_reads(reads[:30] + reads[-4:])
    _check_against(fz, dz, bz, ef.find_sites(bz, motif_seqs, 0, chrom2len, K), K, S, False, "zscore")
    assert len(feats) == len(sites) > 300
    lens = np.array([f[9] for f in feats])
    assert (lens.sum(1) > 448).sum() >= 5 and (lens.sum(1) <= 448).sum() > 200     # windowed and recomputing sites
    for round_stats in (False, True):
        got = _check_against(feats, drawn, batch, sites, K, S, round_stats)
    assert got["read_scale"][-3].item() == 0.0


@pytest.mark.gpu
def test_gpu_philox_subsample_is_an_ordered_uniform_subset():
    K, S = 13, 16
    reads = synthetic.make_reads(40, seed=7, mean_bases=200, mean_dwell=14.0, long_every=2)
    motif_seqs = eo.get_motif_seqs("CG")
    batch = ef.pack_reads(reads)
    sites = ef.find_sites(batch, motif_seqs, 0, None, K)
    feats, drawn = eo.extract_features(reads, "mad", motif_seqs, 0, None, K, S, 1, rng=random.Random(3))
    want = eo.features_to_arrays(feats, round_stats=False)
    a = ef.extract_tensors(batch, sites, K, S, seed=5)
    b = ef.extract_tensors(batch, sites, K, S, seed=5)
    c = ef.extract_tensors(batch, sites, K, S, seed=6)
    sa, sb, sc = (t["signals"].cpu().numpy() for t in (a, b, c))
    lens = want["base_signal_lens"]
    short = lens <= S
    assert np.array_equal(sa, sb)                                        # deterministic in the seed
    assert np.array_equal(sa[short], want["signals"][short])             # nothing random about short bases
    assert not np.array_equal(sa[~short], sc[~short])                    # another seed, another draw
    for k in ("kmer", "base_means", "base_stds", "base_signal_lens"):
        assert np.array_equal(a[k].cpu().numpy(), want[k])
    # every long base: the row is a subsequence of the base's normalised samples (strictly increasing offsets)
    norm = {}
    picks, frac = 0, []
    for i, f in enumerate(feats):
        r = int(sites.site_read[i])
        if r not in norm:
            rd = reads[r]
            x = rd["raw"] if rd["scaling"] is None else eo.rescale_signals(rd["raw"], rd["scaling"], rd["offset"])
            norm[r] = eo.normalize_signals(x)
        for j in range(K):
            n = int(lens[i, j])
            if n <= S:
                continue
            ev = int(sites.site_ev[i]) - (K - 1) // 2 + j
            base = norm[r][batch.ev_start[ev]:batch.ev_start[ev] + n].astype(np.float32)
            pos = -1
            for v in sa[i, j]:
                nxt = np.nonzero(base[pos + 1:] == v)[0]
                assert nxt.size, (i, j)
                pos = pos + 1 + int(nxt[0])
                frac.append(pos / (n - 1))
            picks += 1
    assert picks > 100
    # offsets of a uniform subset are uniform over the base: mean 1/2, and every decile is populated evenly
    frac = np.array(frac)
    assert abs(frac.mean() - 0.5) < 0.02
    hist = np.histogram(frac, bins=10, range=(0, 1))[0] / frac.size
    assert np.abs(hist - 0.1).max() < 0.03


@pytest.mark.gpu
def test_gpu_extracted_tensors_feed_the_classifier_like_oracle_features():
    # extract -> ModelBiLSTM.forward without the feature-file detour == the same model on the oracle's features
    from deepsignal_plant_b200.models import ModelBiLSTM
    K, S = 13, 16
    reads = synthetic.make_reads(80, seed=21, mean_bases=300)
    motif_seqs = eo.get_motif_seqs("CG")
    batch = ef.pack_reads(reads)
    sites = ef.find_sites(batch, motif_seqs, 0, None, K)
    feats, drawn = eo.extract_features(reads, "mad", motif_seqs, 0, None, K, S, 1, rng=random.Random(9))
    want = eo.features_to_arrays(feats, round_stats=False)
    got = ef.extract_tensors(batch, sites, K, S, drawn=eo.drawn_to_array(drawn, K, S))
    torch.manual_seed(1234)
    model = ModelBiLSTM(K, S, seed=77).cuda().eval()
    names = ("kmer", "base_means", "base_stds", "base_signal_lens", "signals")
    with torch.no_grad():
        model._calls = 0
        l1, p1 = model(*[got[k] for k in names])
        model._calls = 0
        l2, p2 = model(*[torch.from_numpy(want[k]).cuda() for k in names])
    assert len(sites) > 500 and torch.equal(p1, p2) and torch.equal(l1, l2)
    f2, err = ef._extract_features(reads, "mad", motif_seqs, 0, None, K, S, 1, None, (None, None, None),
                                   drawn=eo.drawn_to_array(drawn, K, S))
    assert err == 0 and [f[:7] for f in f2] == [f[:7] for f in feats] and [f[9] for f in f2] == [f[9] for f in feats]
    assert np.array_equal(np.array([f[10] for f in f2], np.float32), want["signals"])


@pytest.mark.gpu
def test_gpu_extract_empty_and_invalid():
    reads = synthetic.make_reads(2, seed=3, mean_bases=40)
    batch = ef.pack_reads(reads)
    none = ef.find_sites(batch, ["ACGTACGTACGTAAAA"], 0, None, 13)
    assert len(none) == 0
    out = ef.extract_tensors(batch, none, 13, 16)
    assert out["signals"].shape == (0, 13, 16) and torch.isfinite(out["read_shift"]).all()
    L = _native.lib()
    assert L.dsp_extract_features(0, None, None, None, None, 0, None, None, None, None, None, 0, 12, 16, 0, 0,
                                  None, 0, None, None, None, None, None, None, None, None) == 1
    assert b"odd" in L.dsp_last_error()


@pytest.mark.gpu
def test_gpu_call_mods_from_a_decoded_reads_archive(tmp_path):
    # `call_mods -i reads.npz`: extract + classify in one pass, features never become text
    from deepsignal_plant_b200 import cli, feature_io
    from deepsignal_plant_b200.models import ModelBiLSTM
    from oracle import model_oracle
    K, S = 13, 16
    reads = synthetic.make_reads(45, seed=33, mean_bases=400, long_every=5)
    arch = str(tmp_path / "reads.npz")
    ef.save_reads(arch, reads)
    fasta = str(tmp_path / "genome.fa")
    with open(fasta, "w") as f:
        for c in (1, 2, 3):
            f.write(">chr%d some description\n" % c + ("ACGT" * 25 + "\n") * 2000)
    torch.manual_seed(1234)
    ref = ModelBiLSTM(K, S, 3, 1, 2, 0, 256, 16, 4, True, True)
    ckpt = str(tmp_path / "m.ckpt")
    torch.save(ref.state_dict(), ckpt)
    params = {k: v.detach().numpy() for k, v in ref.state_dict().items()}
    out = str(tmp_path / "calls.tsv")
    argv = ["call_mods", "-i", arch, "-m", ckpt, "-o", out, "--max_batch", "512", "--f5_batch_size", "4",
            "--motifs", "CG", "--reference_path", fasta]
    assert cli.main(argv) == 0
    lines = open(out).read().splitlines()
    motif_seqs = eo.get_motif_seqs("CG")
    feats, _ = eo.extract_features(reads, "mad", motif_seqs, 0, {"chr%d" % c: 200000 for c in (1, 2, 3)}, K, S, 1,
                                   rng=random.Random(0))
    n = len(feats)
    assert len(lines) == n > 1000
    arr = eo.features_to_arrays(feats, round_stats=False)
    cfg = model_oracle.make_cfg()
    zeros = {g: (np.zeros((l * 2, n, h), np.float32),) * 2 for g, l, h in (("seq", 1, 128), ("signal", 1, 128), ("comb", 3, 256))}
    want = model_oracle.forward(params, cfg, *(arr[k] for k in cases.FEATURE_KEYS), zeros)[1]
    for i, line in enumerate(lines):
        w = line.split("\t")
        f = feats[i]
        assert len(w) == 10 and w[:6] == [f[0], str(f[1]), f[2], str(f[3]), f[4], f[5]]
        p0, p1 = float(w[6]), float(w[7])
        assert abs(p0 + p1 - 1.0) < 2e-6 and abs(p1 - want[i, 1]) < 0.05
        assert w[8] == ("1" if p1 > p0 else "0") or abs(p1 - p0) < 2e-6
        assert w[9] == f[6][4:9]
    # region + positions filters reach the extraction
    some = feats[n // 2]
    pos_file = str(tmp_path / "pos.tsv")
    with open(pos_file, "w") as pf:
        for f in feats[::2]:
            pf.write("%s\t%d\t%s\n" % (f[0], f[1], f[2]))
    out2 = str(tmp_path / "calls2.tsv")
    assert cli.main(argv[:6] + [out2] + argv[7:] + ["--positions", pos_file, "--region", some[0]]) == 0
    got = ["\t".join(l.split("\t")[:3]) for l in open(out2).read().splitlines()]
    assert got == ["\t".join([f[0], str(f[1]), f[2]]) for f in feats[::2] if f[0] == some[0]] and len(got) > 50


@pytest.mark.gpu
def test_gpu_find_sites_matches_host_search_and_reference():
    for name in EXTRACT_CASES:
        z, reads, K, S, chrom2len, motif_seqs, mod_loc = load(name)
        batch = ef.pack_reads(reads)
        dev = ef.find_sites_device(batch, motif_seqs, mod_loc, chrom2len, K)
        assert ef.sampleinfo(batch, dev) == list(z["info"])
    reads = synthetic.make_reads(40, seed=12, mean_bases=300)
    chrom2len = {"chr1": 200000, "chr2": 200000}
    batch = ef.pack_reads(reads)
    full = ef.find_sites(batch, eo.get_motif_seqs("CHG"), 0, chrom2len, 13)
    some = (batch.chrom[int(full.site_read[len(full) // 2])], int(full.pos[len(full) // 2]))
    regions = [(None, None, None), ("chr2", None, None), (some[0], some[1] - 40, some[1] + 25), ("chrNone", None, None),
               (some[0], some[1], None)]
    for motifs, mod_loc, K in (("CHG", 0, 13), ("CHG", 2, 9), ("CG", 1, 5), ("N", 0, 3), ("CHH,CHG", 0, 17)):
        ms = eo.get_motif_seqs(motifs)
        for region in regions:
            host = ef.find_sites(batch, ms, mod_loc, chrom2len, K, regioninfo=region)
            dev = ef.find_sites_device(batch, ms, mod_loc, chrom2len, K, regioninfo=region)
            for f in ("site_read", "site_ev", "pos", "pos_in_strand"):
                assert np.array_equal(getattr(host, f), getattr(dev, f)), (motifs, region, f)
            assert getattr(dev, f).dtype == getattr(host, f).dtype
    assert len(ef.find_sites_device(batch, eo.get_motif_seqs("N"), 0, None, 3)) > batch.ev_base.shape[0] // 2   # capacity retry
    positions = {"||".join([batch.chrom[r], str(int(p)), batch.alignstrand[r]]) for r, p in zip(full.site_read[::3], full.pos[::3])}
    sub = ef.find_sites_device(batch, eo.get_motif_seqs("CHG"), 0, chrom2len, 13, positions=positions)
    assert len(sub) == len(positions)
    # the device-found sites feed the extraction without another upload, same tensors as from the host search
    a = ef.extract_tensors(batch, full, 13, 16, seed=3)
    b = ef.extract_tensors(batch, ef.find_sites_device(batch, eo.get_motif_seqs("CHG"), 0, chrom2len, 13), 13, 16, seed=3)
    assert all(torch.equal(a[k], b[k]) for k in a)


@pytest.mark.gpu
def test_gpu_extract_at_scale_is_batching_invariant_and_matches_oracle_sample():
    # ~90 k sites: whole batch == two halves (bit for bit: nothing depends on what else is in the batch),
    # zero padding is centred, lens are the event lengths, and a sample of reads equals the oracle
    K, S = 13, 16
    base = synthetic.make_reads(120, seed=55, mean_bases=3000, long_every=5, stall_every=40)
    reads = [dict(base[i % len(base)], readname="r%05d" % i) for i in range(480)]
    ms = eo.get_motif_seqs("CG")
    whole = ef.pack_reads(reads)
    sw = ef.find_sites_device(whole, ms, 0, None, K)
    tw = ef.extract_tensors(whole, sw, K, S, seed=1)
    assert len(sw) > 80000
    parts = [ef.pack_reads(reads[:200]), ef.pack_reads(reads[200:])]
    tp = [ef.extract_tensors(b, ef.find_sites_device(b, ms, 0, None, K), K, S, seed=1) for b in parts]
    lens = tw["base_signal_lens"]
    short = (lens <= S)
    for k in ("kmer", "base_means", "base_stds", "base_signal_lens"):
        assert torch.equal(tw[k], torch.cat([t[k] for t in tp])), k
    sig_parts = torch.cat([t["signals"] for t in tp])
    assert torch.equal(tw["signals"][short], sig_parts[short])
    assert torch.equal(tw["read_shift"], torch.cat([t["read_shift"] for t in tp]))
    assert torch.equal(tw["read_scale"], torch.cat([t["read_scale"] for t in tp]))
    # structure of the rectangle: a short base occupies a centred run of n columns, zeros outside
    ev = torch.from_numpy(sw.site_ev).cuda()[:, None] + torch.arange(-(K // 2), K // 2 + 1, device="cuda")[None, :]
    assert torch.equal(lens.long(), torch.from_numpy(whole.ev_len).cuda()[ev])
    col = torch.arange(S, device="cuda")[None, None, :]
    left = ((S - lens.long()) // 2)[:, :, None]
    outside = (col < left) | (col >= left + lens.long()[:, :, None])
    assert (tw["signals"][short][outside[short]] == 0).all()
    # a sample of reads against the oracle (same sites, stats bit for bit, short rows bit for bit)
    feats, _ = eo.extract_features(reads[:6], "mad", ms, 0, None, K, S, 1, rng=random.Random(0))
    want = eo.features_to_arrays(feats, round_stats=False)
    n = len(feats)
    assert np.array_equal(sw.site_read[:n], np.repeat(np.arange(6), np.bincount(sw.site_read[:n], minlength=6)))
    for k in ("kmer", "base_means", "base_stds", "base_signal_lens"):
        assert np.array_equal(tw[k][:n].cpu().numpy().view(np.uint32), want[k].view(np.uint32)), k
    sh = short[:n].cpu().numpy()
    assert np.array_equal(tw["signals"][:n].cpu().numpy()[sh], want["signals"][sh])


@pytest.mark.gpu
@pytest.mark.parametrize("name", EXTRACT_CASES)
def test_gpu_feature_file_lines_equal_the_reference_bytes(name):
    # float64 outputs of the kernels + native formatter == what the reference's _features_to_str wrote
    z, reads, K, S, chrom2len, motif_seqs, mod_loc = load(name)
    batch = ef.pack_reads(reads)
    sites = ef.find_sites_device(batch, motif_seqs, mod_loc, chrom2len, K)
    t = ef.extract_tensors(batch, sites, K, S, normalize_method=method_of(z), round_stats=True, drawn=z["drawn"],
                           dtype=torch.float64)
    assert np.array_equal(t["base_means"].cpu().numpy(), np.around(z["means"], 6))      # float64, bit for bit
    assert np.array_equal(t["base_stds"].cpu().numpy(), np.around(z["stds"], 6))
    assert np.array_equal(t["signals"].cpu().numpy(), z["rect"])
    host = [t[k].cpu().numpy() for k in ("base_means", "base_stds", "base_signal_lens", "signals")]
    text = ef.format_features(batch, sites, *host, 1)
    assert text.decode().splitlines() == [str(x) for x in z["lines"]]


@pytest.mark.gpu
def test_gpu_extract_cli_writes_the_reference_feature_file_format(tmp_path):
    import gzip
    from deepsignal_plant_b200 import cli, feature_io
    from oracle import features_oracle
    K, S = 13, 16
    reads = synthetic.make_reads(40, seed=61, mean_bases=300, long_every=6)
    arch = str(tmp_path / "reads.npz")
    ef.save_reads(arch, reads)
    out = str(tmp_path / "features.tsv")
    assert cli.main(["extract", "-i", arch, "-o", out, "--motifs", "CG", "--f5_batch_size", "7", "--gzip", "--methy_label", "0"]) == 0
    lines = gzip.open(out + ".gz", "rt").read().splitlines()
    feats, _ = eo.extract_features(reads, "mad", eo.get_motif_seqs("CG"), 0, None, K, S, 0, rng=random.Random(0))
    want = eo.features_to_arrays(feats, round_stats=True)
    info, kmers, means, stds, lens, sig, labels = features_oracle.read_features(lines)
    assert len(lines) == len(feats) > 500 and set(labels) == {0}
    assert info == ["\t".join([f[0], str(f[1]), f[2], str(f[3]), f[4], f[5]]) for f in feats]
    assert np.array_equal(np.asarray(kmers, np.float32), want["kmer"])
    assert np.array_equal(np.asarray(means, np.float32), want["base_means"])
    assert np.array_equal(np.asarray(stds, np.float32), want["base_stds"])
    assert np.array_equal(np.asarray(lens, np.float32), want["base_signal_lens"])
    short = want["base_signal_lens"] <= S
    assert np.array_equal(np.asarray(sig, np.float32)[short], want["signals"][short])
    # and the file is what this repository's own reader / call_mods take
    rd = feature_io.FeatureFileReader(out + ".gz", K, S, batch_sites=4096, slots=3, nthreads=2)
    assert sum(b.n for b in rd) == len(lines)


def test_one_bad_read_is_skipped_not_fatal(tmp_path):
    # the reference counts a failing read and goes on (extract_features.py:373-375)
    from deepsignal_plant_b200 import extract_features as ef
    from deepsignal_plant_b200 import synthetic
    reads = synthetic.make_reads(6, seed=3, mean_bases=60)
    good = ef.pack_reads(reads)
    arrays = good.arrays()
    lo, hi = int(good.ev_off[2]), int(good.ev_off[3])
    arrays["ev_base"] = arrays["ev_base"].copy()
    arrays["ev_base"][lo + 5] = ord("x")                                  # a letter outside base2code_dna
    arrays["ev_len"] = arrays["ev_len"].copy()
    arrays["ev_len"][int(good.ev_off[5]) - 1] += 10 ** 6                   # last event of read 4 runs off its raw signal
    path = str(tmp_path / "reads.npz")
    np.savez(path, **arrays)
    got = ef.load_reads(path)
    assert got.n_errors == 2 and got.n_reads == 4
    assert list(got.readname) == [good.readname[i] for i in (0, 1, 3, 5)]
    assert not got.bad_reads().any()
    want = ef.pack_reads([reads[i] for i in (0, 1, 3, 5)])
    for f in ("raw", "raw_off", "ev_off", "ev_start", "ev_len", "ev_base", "chrom_start"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f


def test_mad_constant_equals_scipy_norm_ppf_statsmodels_itself_absent_here():
    # statsmodels.robust.mad(a) = median(|a - median(a)| / c) with c = scipy.stats.norm.ppf(3/4) (statsmodels/robust/scale.py,
    # `Gaussian = scipy.stats.norm`).  statsmodels is not in this image, so the formula stays "parity unpinned" against it
    # (DESIGN.md section 1); the constant is pinned here against scipy, the library statsmodels takes it from, and against
    # statsmodels itself wherever that can be imported.
    from scipy.stats import norm
    from oracle import extract_oracle as eo
    assert eo.MAD_C == float(norm.ppf(0.75))
    try:
        from statsmodels import robust
    except Exception:
        return
    if getattr(robust, "__dsp_stub__", False) or not hasattr(robust, "mad") or robust.mad is None:
        return
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 1000, 1001):
        a = rng.normal(0, 1, n) * 50
        assert eo.mad(a) == float(robust.mad(a))
