"""CPU: host-side mirror of the reference interface (constructor, state_dict contract,
text formatting, synthetic data) -- no kernel calls."""
import os

import numpy as np
import pytest
import torch

import cases
from deepsignal_plant_b200 import call_modifications as cm
from deepsignal_plant_b200 import synthetic
from deepsignal_plant_b200.models import ModelBiLSTM


def test_constructor_contract():
    m = ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True, module="both_bilstm", device=0)
    assert m.get_model_type() == "BiLSTM"
    sd = m.state_dict()
    assert sum(v.numel() for v in sd.values()) == 4694082
    assert tuple(sd["embed.weight"].shape) == (16, 4)
    assert tuple(sd["lstm_seq.weight_ih_l0"].shape) == (512, 7)
    assert tuple(sd["lstm_signal.weight_ih_l0_reverse"].shape) == (512, 16)
    assert tuple(sd["lstm_comb.weight_ih_l2"].shape) == (1024, 512)
    assert tuple(sd["fc1.weight"].shape) == (256, 512) and tuple(sd["fc2.weight"].shape) == (2, 256)
    with pytest.raises(ValueError, match="--model_type is not right!"):
        ModelBiLSTM(module="bogus")
    seq = ModelBiLSTM(module="seq_bilstm")
    assert "lstm_signal.weight_ih_l0" not in seq.state_dict() and tuple(seq.state_dict()["lstm_seq.weight_hh_l0"].shape) == (1024, 256)
    sig = ModelBiLSTM(module="signal_bilstm")
    assert "embed.weight" not in sig.state_dict()


def test_checkpoint_roundtrip_like_call_mods_q(tmp_path):
    # call_modifications.py:219-223: torch.load -> dict.update -> load_state_dict
    torch.manual_seed(3)
    a = ModelBiLSTM(hidden_size=32)
    path = tmp_path / "both_bilstm.b13_s16_epoch1.ckpt"
    torch.save(a.state_dict(), path)
    b = ModelBiLSTM(hidden_size=32)
    para = torch.load(path, map_location=torch.device("cpu"))
    d = b.state_dict()
    d.update(para)
    b.load_state_dict(d)
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k])


def test_train_mode_is_refused():
    m = ModelBiLSTM(hidden_size=32)
    m.train()
    x = torch.zeros(2, 13)
    with pytest.raises(RuntimeError, match="inference only"):
        m(x, x, x, x, torch.zeros(2, 13, 16))


def test_init_hidden_consumes_rng_like_reference():
    m = ModelBiLSTM(hidden_size=32)
    torch.manual_seed(5)
    h0, c0 = m.init_hidden(7, 3, 32)
    torch.manual_seed(5)
    assert torch.equal(h0, torch.randn(6, 7, 32)) and torch.equal(c0, torch.randn(6, 7, 32))


def test_format_calls_matches_reference_lines():
    e = cases.MANIFEST["callmods"]
    probs = np.load(cases.GOLD + "/callmods_%d_probs.npz" % e["rng_seed"])["probs"]
    lines = cases.read_gz("callmods_%d.tsv.gz" % e["rng_seed"]).splitlines()
    feats = synthetic.make_features(e["n"], 13, 16, seed=e["feature_seed"])
    info = synthetic.make_sampleinfo(e["n"], seed=e["feature_seed"])
    got = cm.format_calls(info, feats["kmer"], probs, probs.argmax(1))
    assert got == lines


def test_format_calls_extreme_probabilities():
    # scientific notation, exact 0/1, ties: must print like str(np.float32) does
    probs = np.array([[1e-6, 1 - 1e-6], [0.0, 1.0], [1.0, 0.0], [0.5, 0.5], [5.6e-5, 1 - 5.6e-5],
                      [0.3333333, 0.6666667]], np.float32)
    from oracle import callmods_oracle
    info = ["c\t1\t+\t1\tr\tt"] * len(probs)
    kmers = np.tile(np.arange(13) % 16, (len(probs), 1))
    want, labels = callmods_oracle.call_lines(info, kmers, probs)
    assert cm.format_calls(info, kmers, probs, labels) == want
    assert want[0].split("\t")[6] == "1e-06" and want[1].split("\t")[6:8] == ["0.0", "1.0"]


def test_kmer_centre_short_kmers():
    assert cm.kmer_centre(np.array([[0, 1, 2]])).tolist() == ["ACG"]
    assert cm.kmer_centre(np.array([[0, 1, 2, 3, 4, 5, 6]], np.float32)).tolist() == ["CGTNW"]


def test_synthetic_rectangle_rule():
    f = synthetic.make_features(64, 13, 16, seed=3)
    lens = f["base_signal_lens"].astype(int)
    sig = f["signals"]
    for n in range(64):
        for t in range(13):
            L = lens[n, t]
            row = sig[n, t]
            if L < 16:
                left = (16 - L) // 2
                assert (row[:left] == 0).all() and (row[left + L:] == 0).all()
            assert np.count_nonzero(row) <= min(L, 16)
    assert (f["kmer"][:, 6] == 1).all()
    g = synthetic.make_features(64, 13, 16, seed=3)
    assert all(np.array_equal(f[k], g[k]) for k in f)


def test_command_line_accepts_the_reference_flags():
    # every --flag of the reference's extract / call_mods / call_freq parsers (deepsignal_plant.py:120-316, 438-475,
    # listed here so that the test runs without /root/reference) parses here too, with the reference's defaults
    from deepsignal_plant_b200 import cli
    ref = {
        "extract": ["--fast5_dir", "--recursively", "--corrected_group", "--basecall_subgroup", "--is_dna", "--reference_path",
                    "--normalize_method", "--methy_label", "--seq_len", "--signal_len", "--motifs", "--mod_loc", "--region",
                    "--positions", "--write_path", "--w_is_dir", "--w_batch_num", "--gzip", "--nproc", "--f5_batch_size"],
        "call_mods": ["--input_path", "--f5_batch_size", "--model_path", "--model_type", "--seq_len", "--signal_len", "--layernum1",
                      "--layernum2", "--class_num", "--dropout_rate", "--n_vocab", "--n_embed", "--is_base", "--is_signallen",
                      "--batch_size", "--hid_rnn", "--result_file", "--gzip", "--recursively", "--corrected_group",
                      "--basecall_subgroup", "--reference_path", "--is_dna", "--normalize_method", "--methy_label", "--motifs",
                      "--mod_loc", "--region", "--positions", "--nproc", "--nproc_gpu"],
        "call_freq": ["--input_path", "--file_uid", "--result_file", "--bed", "--sort", "--gzip", "--prob_cf", "--contigs", "--nproc"],
    }
    parser = cli.build_parser()
    sub = [a for a in parser._actions if a.__class__.__name__ == "_SubParsersAction"][0]
    for name, flags in ref.items():
        have = {o for a in sub.choices[name]._actions for o in a.option_strings}
        assert not [f for f in flags if f not in have], name
    a = parser.parse_args(["extract", "-i", "reads.npz", "-o", "f.tsv"])
    assert (a.normalize_method, a.methy_label, a.seq_len, a.signal_len, a.motifs, a.mod_loc, a.f5_batch_size, a.nproc, a.w_is_dir) == \
        ("mad", 1, 13, 16, "CG", 0, 30, 10, "no")
    c = parser.parse_args(["call_mods", "-i", "x", "-m", "m.ckpt", "-o", "o.tsv"])
    assert (c.model_type, c.seq_len, c.signal_len, c.layernum1, c.layernum2, c.class_num, c.hid_rnn, c.n_vocab, c.n_embed,
            c.batch_size, c.motifs, c.mod_loc) == ("both_bilstm", 13, 16, 3, 1, 2, 256, 16, 4, 512, "CG", 0)
    f = parser.parse_args(["call_freq", "-i", "a", "-i", "b", "-o", "o"])
    assert f.input_path == ["a", "b"] and f.prob_cf == 0.5 and not f.bed and not f.sort


def test_rank_parts_merge_in_rank_order(tmp_path):
    # call_mods under torchrun: every rank copies its own part to its offset of the one result file (gloo, 3 ranks)
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "calls.tsv")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "3", "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + os.getpid() % 500), os.path.join(root, "tests", "merge_parts_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    want = b"".join((b"rank %d line\n" % k) * (0 if k == 1 else 1000 * (k + 1) + 7) for k in range(3))
    assert open(out, "rb").read() == want
    assert sorted(os.listdir(tmp_path)) == ["calls.tsv"]               # the parts are gone
