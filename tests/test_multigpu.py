"""Multi-GPU paths on a box with >= 2 B200s (skipped otherwise): the forward shards by site
batch with no collective (bench.py under torchrun), call_freq adds one exchange over NVLink peer memory (csrc/comm.cu) and must
stay bit-exact."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, script, *args, port=29533):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:]
    return json.loads(lines[-1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_freq_exchange_over_nvlink_is_bit_exact():
    out = _torchrun(2, "tools/bench_freq_dist.py", "--records_per_rank", "2000000", "--iters", "2")
    assert out["bit_exact"] is True and out["world"] == 2 and out["coverage_sum_equals_callable"] and out["slices_ordered"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bench_two_ranks_weak_scaling_line():
    out = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "20", "--warmup", "3", port=29534)
    assert out["n_gpus"] == 2 and out["scaling"] == "weak" and out["value"] > 0 and out["gpu_launches"] > 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_call_mods_cli_two_ranks_keeps_file_order(tmp_path):
    import numpy as np
    from deepsignal_plant_b200 import feature_io, synthetic
    from deepsignal_plant_b200.models import ModelBiLSTM
    n = 5000
    feats = synthetic.make_features(n, 13, 16, seed=43)
    info = synthetic.make_sampleinfo(n, seed=43)
    path = str(tmp_path / "features.tsv")
    with open(path, "w") as f:
        for i in range(n):
            f.write(feature_io.features_to_str(info[i], feats["kmer"][i], feats["base_means"][i], feats["base_stds"][i],
                                               feats["base_signal_lens"][i], feats["signals"][i], 0) + "\n")
    torch.manual_seed(1234)
    ckpt = str(tmp_path / "m.ckpt")
    torch.save(ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True).state_dict(), ckpt)
    out = str(tmp_path / "calls.tsv")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29535", "-m", "deepsignal_plant_b200", "call_mods", "-i", path, "-m", ckpt, "-o", out,
           "--max_batch", "1024"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = open(out).read().splitlines()
    assert len(lines) == n and ["\t".join(l.split("\t")[:6]) for l in lines] == info
    assert not [p for p in os.listdir(tmp_path) if ".part" in p]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_call_mods_archive_two_ranks_keeps_read_order(tmp_path):
    # decoded-reads archive sharded by contiguous read ranges, one process per GPU, no collective on the data path
    import random
    from deepsignal_plant_b200 import extract_features as ef, synthetic
    from deepsignal_plant_b200.models import ModelBiLSTM
    from oracle import extract_oracle as eo
    reads = synthetic.make_reads(31, seed=44, mean_bases=300)
    arch = str(tmp_path / "reads.npz")
    ef.save_reads(arch, reads)
    torch.manual_seed(1234)
    ckpt = str(tmp_path / "m.ckpt")
    torch.save(ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True).state_dict(), ckpt)
    out = str(tmp_path / "calls.tsv")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29536", "-m", "deepsignal_plant_b200", "call_mods", "-i", arch, "-m", ckpt, "-o", out,
           "--max_batch", "1024", "--f5_batch_size", "5", "--motifs", "CG"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    feats, _ = eo.extract_features(reads, "mad", eo.get_motif_seqs("CG"), 0, None, 13, 16, 1, rng=random.Random(0))
    lines = open(out).read().splitlines()
    assert len(lines) == len(feats) > 400
    assert [l.split("\t")[:6] for l in lines] == [[f[0], str(f[1]), f[2], str(f[3]), f[4], f[5]] for f in feats]
    assert [l.split("\t")[9] for l in lines] == [f[6][4:9] for f in feats]
    assert not [p for p in os.listdir(tmp_path) if ".part" in p]
