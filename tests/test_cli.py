"""call_mods / call_freq command line end to end on the GPU: feature file -> native parser ->
pinned batches -> CUDA forward -> native formatter -> output file -> frequency table."""
import gzip
import os

import numpy as np
import pytest
import torch

import cases
from deepsignal_plant_b200 import cli, feature_io, synthetic
from deepsignal_plant_b200 import call_mods_freq as cf
from oracle import model_oracle, freq_oracle

pytestmark = pytest.mark.gpu


def write_inputs(tmp_path, n, seed=41, gz=False):
    feats = synthetic.make_features(n, 13, 16, seed=seed)
    info = synthetic.make_sampleinfo(n, seed=seed)
    labels = np.random.default_rng(seed).integers(0, 2, n)
    path = str(tmp_path / ("features.tsv.gz" if gz else "features.tsv"))
    opener = gzip.open if gz else open
    with opener(path, "wt") as f:
        for i in range(n):
            f.write(feature_io.features_to_str(info[i], feats["kmer"][i], feats["base_means"][i], feats["base_stds"][i],
                                               feats["base_signal_lens"][i], feats["signals"][i], labels[i]) + "\n")
    torch.manual_seed(1234)
    from deepsignal_plant_b200.models import ModelBiLSTM
    ref = ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True)
    ckpt = str(tmp_path / "both_bilstm.b13_s16_epoch1.ckpt")
    torch.save(ref.state_dict(), ckpt)                      # train.py:161-164: a bare state_dict
    return path, ckpt, feats, info, {k: v.detach().numpy() for k, v in ref.state_dict().items()}


@pytest.mark.parametrize("gz", [False, True])
def test_call_mods_cli_end_to_end(tmp_path, gz):
    n = 3001
    path, ckpt, feats, info, params = write_inputs(tmp_path, n, gz=gz)
    out = str(tmp_path / "calls.tsv")
    argv = ["call_mods", "-i", path, "-m", ckpt, "-o", out, "--max_batch", "1024", "--nproc", "4"] + (["--gzip"] if gz else [])
    assert cli.main(argv) == 0
    text = (gzip.open(out + ".gz", "rt") if gz else open(out)).read()
    lines = text.splitlines()
    assert len(lines) == n
    # the oracle with ZERO initial states brackets what the Philox-state run may produce
    cfg = model_oracle.make_cfg()
    zeros = {g: (np.zeros((l * 2, n, h), np.float32),) * 2 for g, l, h in (("seq", 1, 128), ("signal", 1, 128), ("comb", 3, 256))}
    want = model_oracle.forward(params, cfg, *(feats[k] for k in cases.FEATURE_KEYS), zeros)[1]
    for i, line in enumerate(lines):
        w = line.split("\t")
        assert len(w) == 10 and "\t".join(w[:6]) == info[i]
        p0, p1 = float(w[6]), float(w[7])
        assert abs(p0 + p1 - 1.0) < 2e-6 and abs(p1 - want[i, 1]) < 0.05
        assert w[8] == ("1" if p1 > p0 else "0") or abs(p1 - p0) < 2e-6
        assert w[9] == "".join(feature_io.code2base_dna[int(c)] for c in feats["kmer"][i][4:9])
    # call_freq on that output, against the oracle on the same lines
    freq = str(tmp_path / "freq.tsv")
    assert cli.main(["call_freq", "-i", out + (".gz" if gz else ""), "-o", freq, "--prob_cf", "0.0", "--sort"]) == 0
    assert open(freq).read() == freq_oracle.render(freq_oracle.aggregate(lines, 0.0), True, False)


def test_call_mods_cli_rejects_fast5_directory(tmp_path):
    path, ckpt, *_ = write_inputs(tmp_path, 10)
    with pytest.raises(ValueError, match="fast5"):
        cli.main(["call_mods", "-i", str(tmp_path), "-m", ckpt, "-o", str(tmp_path / "o.tsv")])
    with pytest.raises(ValueError, match="model_path"):
        cli.main(["call_mods", "-i", path, "-m", str(tmp_path / "missing.ckpt"), "-o", str(tmp_path / "o.tsv")])


def test_reference_worker_protocol(tmp_path):
    # _read_features_file -> _call_mods_q -> _write_predstr_to_file with queues (call_modifications.py:587-626)
    from deepsignal_plant_b200 import call_modifications as cm
    n = 700
    path, ckpt, feats, info, params = write_inputs(tmp_path, n)
    args = cli.build_parser().parse_args(["call_mods", "-i", path, "-m", ckpt, "-o", str(tmp_path / "o.tsv"), "-b", "256"])
    fq, pq = cm.SimpleQueue(), cm.SimpleQueue()
    cm._read_features_file(path, fq, 3)
    cm._call_mods_q(ckpt, fq, pq, None, args, 0)
    pq.put("kill")
    cm._write_predstr_to_file(str(tmp_path / "o.tsv"), pq, False)
    lines = open(tmp_path / "o.tsv").read().splitlines()
    assert len(lines) == n and ["\t".join(l.split("\t")[:6]) for l in lines] == info
    assert fq.get() == "kill"


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2])
def test_call_mods_freq_out_equals_call_freq_on_the_written_file(tmp_path, world):
    # call_mods --freq_out: the per-site table of the run's own calls, without re-reading the calls file; with 2 ranks
    # (sharing the GPU here) the ranks' calls meet through the NVLink-style exchange.  Must equal call_freq on the file.
    import subprocess
    import sys
    from deepsignal_plant_b200 import feature_io, synthetic
    from deepsignal_plant_b200.models import ModelBiLSTM
    from oracle import freq_oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = 6000
    feats = synthetic.make_features(n, 13, 16, seed=51)
    info = synthetic.make_sampleinfo(n, seed=51, n_chrom=3, n_pos=300)
    path = str(tmp_path / "features.tsv")
    with open(path, "w") as f:
        for i in range(n):
            f.write(feature_io.features_to_str(info[i], feats["kmer"][i], feats["base_means"][i], feats["base_stds"][i],
                                               feats["base_signal_lens"][i], feats["signals"][i], 0) + "\n")
    torch.manual_seed(1234)
    ckpt = str(tmp_path / "m.ckpt")
    torch.save(ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True).state_dict(), ckpt)
    out, freq = str(tmp_path / "calls.tsv"), str(tmp_path / "freq.tsv")
    tail = ["-m", "deepsignal_plant_b200", "call_mods", "-i", path, "-m", ckpt, "-o", out, "--max_batch", "1024",
            "--freq_out", freq, "--freq_prob_cf", "0.002", "--freq_sort"]
    cmd = [sys.executable] + tail if world == 1 else \
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
         "--master-port", "29737"] + tail
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = open(out).read().splitlines()
    assert len(lines) == n
    want = freq_oracle.render(freq_oracle.aggregate(lines, 0.002), True, False)
    assert len(want) > 1000 and open(freq).read() == want
