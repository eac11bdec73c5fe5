"""CPU: the oracle restatements against the fixtures the reference produced."""
import numpy as np
import pytest

import cases
from deepsignal_plant_b200 import synthetic
from oracle import model_oracle, callmods_oracle, freq_oracle


@pytest.mark.parametrize("name", cases.FORWARD_CASES)
def test_model_oracle_matches_reference_golden(name):
    case = cases.load_case(name)                     # also checks the sha256 pins
    case = cases.slice_case(case, min(case["entry"]["n"], 1024))
    f = case["feats"]
    logits, probs = model_oracle.forward(case["params"], case["cfg"], f["kmer"], f["base_means"], f["base_stds"],
                                         f["base_signal_lens"], f["signals"], case["states"])
    assert np.abs(probs - case["probs"]).max() < 5e-6
    assert np.abs(logits - case["logits"]).max() < 2e-5
    assert (probs.argmax(1) == case["probs"].argmax(1)).all()


def test_flops_per_site_match_survey():
    # SURVEY.md section 8d
    assert model_oracle.flops_per_site(model_oracle.make_cfg()) == 118447104
    assert model_oracle.flops_per_site(model_oracle.make_cfg(module="seq_bilstm")) == 126727168
    assert model_oracle.flops_per_site(model_oracle.make_cfg(module="signal_bilstm")) == 127206400
    assert model_oracle.flops_per_site(model_oracle.make_cfg(seq_len=17, signal_len=20)) == 154950656


def test_callmods_oracle_matches_reference_lines():
    e = cases.MANIFEST["callmods"]
    probs = np.load(cases.GOLD + "/callmods_%d_probs.npz" % e["rng_seed"])["probs"]
    lines = cases.read_gz("callmods_%d.tsv.gz" % e["rng_seed"]).splitlines()
    feats = synthetic.make_features(e["n"], 13, 16, seed=e["feature_seed"])
    info = synthetic.make_sampleinfo(e["n"], seed=e["feature_seed"])
    got, labels = callmods_oracle.call_lines(info, feats["kmer"], probs)
    assert got == lines
    assert (labels == np.array([int(l.split("\t")[8]) for l in lines])).all()


def _freq_inputs():
    edge = open(cases.GOLD + "/freq_edge_input.tsv").read().splitlines()
    synth = synthetic.make_callmods_records(100000, n_chrom=12, n_pos=900, seed=5)
    return {"edge": edge, "synth": synth}


FREQ_CASES = sorted(k for k in cases.MANIFEST["freq"] if k.startswith("freq_"))


@pytest.fixture(scope="module")
def freq_inputs():
    return _freq_inputs()


@pytest.mark.parametrize("name", FREQ_CASES)
def test_freq_oracle_matches_reference_bytes(name, freq_inputs):
    e = cases.MANIFEST["freq"][name]
    table = freq_oracle.aggregate(freq_inputs[e["input"]], e["prob_cf"])
    assert freq_oracle.render(table, e["sort"], e["bed"]) == cases.read_gz(name + ".txt.gz")


def test_features_oracle_matches_reference_reader():
    import gzip
    import os
    from oracle import features_oracle
    lines = gzip.open(os.path.join(cases.GOLD, "features_small.tsv.gz"), "rt").read().splitlines()
    g = np.load(os.path.join(cases.GOLD, "features_small_parsed.npz"))
    info, kmers, means, stds, lens, sig, labels = features_oracle.read_features(lines)
    for got, key, dt in ((kmers, "kmer", np.float32), (means, "base_means", np.float32), (stds, "base_stds", np.float32),
                         (lens, "base_signal_lens", np.float32), (sig, "signals", np.float32), (labels, "labels", np.int32)):
        assert np.asarray(got, dtype=dt).tobytes() == g[key].tobytes(), key
    assert features_oracle.batch_sizes(lines, cases.MANIFEST["features"]["f5_batch_size"]) == g["batch_sizes"].tolist()


@pytest.mark.parametrize("name", ["names", "names_sorted_bed", "fasta"])
def test_freq_oracle_contigs_mode_matches_reference(name):
    m = cases.MANIFEST["freq_contigs"]
    e = m[name]
    lines = synthetic.make_callmods_records(m["input"]["n"], n_chrom=m["input"]["n_chrom"], n_pos=m["input"]["n_pos"], seed=m["input"]["seed"])
    contigs = e["contigs"].split(",") if e["contigs"] else [l[1:].split(" ")[0] for l in m["fasta_text"].splitlines() if l.startswith(">")]
    assert freq_oracle.render_by_contig(lines, contigs, e["prob_cf"], e["sort"], e["bed"]) == cases.read_gz("freq_contigs_%s.txt.gz" % name)


@pytest.mark.parametrize("name", ["both_13_16_s1", "seq_13_16_s1234", "signal_13_16_s1234", "both_small_odd", "seq_nobase", "both_nolen_c3"])
def test_torch_cpu_oracle_matches_reference_fixtures(name):
    # the CPU baseline of bench.py runs this restatement: same torch CPU operators as the reference
    import torch
    from oracle import torch_oracle
    case = cases.load_case(name)
    n = min(case["entry"]["n"], 2048)
    case = cases.slice_case(case, n)
    sd = {k: torch.from_numpy(v) for k, v in case["params"].items()}
    states = {g: tuple(torch.from_numpy(np.ascontiguousarray(x)) for x in hc) for g, hc in case["states"].items()}
    logits, probs = torch_oracle.forward(sd, case["cfg"], *(torch.from_numpy(case["feats"][k]) for k in cases.FEATURE_KEYS), states)
    assert np.abs(probs.numpy() - case["probs"]).max() < 2e-6 and np.abs(logits.numpy() - case["logits"]).max() < 2e-5
