"""Binary feature hand-off (.dspf, feature_bin.py): what a reader gets from it is bit-identical to what the native parser
gets from the reference's text feature file (itself pinned against the reference reader, tests/test_feature_io.py), for
every batch size, block size and rank shard; and (GPU) the calls written from it equal the calls written from text."""
import gzip
import hashlib
import os

import numpy as np
import pytest

import cases
from deepsignal_plant_b200 import feature_bin, feature_io

GOLD = cases.GOLD
FEAT = cases.MANIFEST["features"]
KEYS = ("kmer", "base_means", "base_stds", "base_signal_lens", "signals")
TEXT = os.path.join(GOLD, "features_small.tsv.gz")


def collect(reader):
    got = {k: [] for k in KEYS + ("labels",)}
    info = []
    for b in reader:
        assert 1 <= b.n <= reader.batch_sites
        for k, t in zip(KEYS, b.arrays()):
            got[k].append(t.numpy().copy())
        got["labels"].append(b.labels.numpy().copy())
        info += b.sampleinfo()
    return {k: (np.concatenate(v, 0) if v else np.empty(0)) for k, v in got.items()}, info


@pytest.mark.parametrize("block_sites", [4096, 37])
@pytest.mark.parametrize("batch_sites", [65536, 50, 1])
def test_binary_file_equals_the_parsed_text_file(tmp_path, block_sites, batch_sites):
    g = np.load(os.path.join(GOLD, "features_small_parsed.npz"))
    p = str(tmp_path / "f.dspf")
    assert feature_bin.pack_feature_file(TEXT, p, batch_sites=block_sites, nthreads=2) == FEAT["n"]
    assert feature_bin.is_feature_bin(p) and not feature_bin.is_feature_bin(TEXT)
    T, S, blocks = feature_bin.scan_blocks(p)
    assert (T, S) == (FEAT["seq_len"], FEAT["signal_len"]) and sum(b[1] for b in blocks) == FEAT["n"]
    assert len(blocks) == -(-FEAT["n"] // block_sites)
    rd = feature_bin.FeatureBinReader(p, batch_sites=batch_sites, pinned=False, slots=2, nthreads=3)
    got, info = collect(rd)
    assert len(info) == FEAT["n"] == rd.sites_read == rd.total_sites
    assert hashlib.sha256("\n".join(info).encode()).hexdigest() == FEAT["sampleinfo_sha256"]
    for k in got:
        assert got[k].dtype == g[k].dtype and got[k].shape == g[k].shape
        assert got[k].tobytes() == g[k].tobytes(), k


def test_file_format_is_pinned(tmp_path):
    # the bytes of the packed golden feature file: a change of the layout (or of a parsed value) shows up here, so files
    # that users have written keep being readable
    p = str(tmp_path / "f.dspf")
    feature_bin.pack_feature_file(TEXT, p, nthreads=2)
    data = open(p, "rb").read()
    assert len(data) == 262592 and data[:8] == b"DSPFEAT1" and data[64:72] == b"DSPFBLK1"
    assert hashlib.sha256(data).hexdigest() == "4e64f983ced19dbae89375954ef314c60526c175a6cdfa429c0e63ddc1903297"


def test_pack_features_command_line(tmp_path, capsys):
    from deepsignal_plant_b200 import cli
    a, b = str(tmp_path / "a.dspf"), str(tmp_path / "b.dspf")
    assert cli.main(["pack_features", "-i", TEXT, "-o", a, "--host_threads", "2"]) == 0
    assert "%d sites" % FEAT["n"] in capsys.readouterr().out
    feature_bin.pack_feature_file(TEXT, b, nthreads=2)
    assert open(a, "rb").read() == open(b, "rb").read()


def test_site_range_shards_cover_the_file_once(tmp_path):
    p = str(tmp_path / "f.dspf")
    n = feature_bin.pack_feature_file(TEXT, p, batch_sites=100, nthreads=2)
    for world in (1, 2, 3, 7):
        info = []
        for r in range(world):
            rd = feature_bin.FeatureBinReader(p, 13, 16, batch_sites=64, pinned=False, site_range=(n * r // world, n * (r + 1) // world))
            info += collect(rd)[1]
        assert hashlib.sha256("\n".join(info).encode()).hexdigest() == FEAT["sampleinfo_sha256"], world


def test_writer_takes_scalar_label_and_offset_views(tmp_path):
    rng = np.random.default_rng(3)
    n, T, S = 11, 5, 4
    arrs = [rng.standard_normal((n, T)).astype(np.float32) for _ in range(4)] + [rng.standard_normal((n, T, S)).astype(np.float32)]
    names = ["c%d\t%d\t+\t%d\tr\tt" % (i, i * 7, i) for i in range(n)]
    text = np.frombuffer(("junk" + "".join(names)).encode(), np.uint8)
    off = 4 + np.concatenate([[0], np.cumsum([len(x) for x in names])])          # offsets that do not start at 0
    p = str(tmp_path / "w.dspf")
    with feature_bin.FeatureBinWriter(p, T, S) as w:
        w.write(*arrs, 1, text, off)
        w.write(*[a[:0] for a in arrs], 1, text, off[:1])                         # an empty block is not written
        w.write(*[a[:3] for a in arrs], np.array([0, 1, 0]), text.tobytes(), off[:4])
    rd = feature_bin.FeatureBinReader(p, batch_sites=8, pinned=False)
    got, info = collect(rd)
    assert info == names + names[:3]
    assert got["labels"].tolist() == [1] * n + [0, 1, 0]
    assert np.array_equal(got["signals"], np.concatenate([arrs[4], arrs[4][:3]]))
    with pytest.raises(ValueError):
        with feature_bin.FeatureBinWriter(str(tmp_path / "bad.dspf"), T, S) as w:
            w.write(*[a[:, :3] for a in arrs[:4]], arrs[4], 1, text, off)


def test_wrong_shape_and_damaged_files_fail_loudly(tmp_path):
    p = str(tmp_path / "f.dspf")
    feature_bin.pack_feature_file(TEXT, p, nthreads=2)
    with pytest.raises(ValueError, match="13-mers"):
        feature_bin.FeatureBinReader(p, 17, 20, pinned=False)
    data = open(p, "rb").read()
    q = str(tmp_path / "cut.dspf")
    open(q, "wb").write(data[:len(data) // 2])
    with pytest.raises(ValueError, match="truncated"):
        feature_bin.FeatureBinReader(q, pinned=False)
    open(q, "wb").write(data[:64] + b"XXXXXXXX" + data[72:])
    with pytest.raises(ValueError, match="damaged block header"):
        feature_bin.scan_blocks(q)
    open(q, "wb").write(gzip.open(TEXT, "rb").read())
    with pytest.raises(ValueError, match="bad magic"):
        feature_bin.scan_blocks(q)


@pytest.mark.gpu
def test_call_mods_on_a_binary_file_writes_the_same_calls_as_on_text(tmp_path):
    import torch
    from deepsignal_plant_b200 import cli
    from deepsignal_plant_b200.models import ModelBiLSTM
    text = str(tmp_path / "f.tsv")
    open(text, "wb").write(gzip.open(TEXT, "rb").read())
    binary = str(tmp_path / "f.dspf")
    feature_bin.pack_feature_file(text, binary)
    ckpt = str(tmp_path / "m.ckpt")
    torch.manual_seed(1234)
    torch.save(ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True).state_dict(), ckpt)
    outs = []
    for src in (text, binary):
        out = str(tmp_path / (os.path.basename(src) + ".calls"))
        cli.main(["call_mods", "-i", src, "-m", ckpt, "-o", out])
        outs.append(open(out, "rb").read())
    # initial states come from the in-kernel Philox stream keyed by (call number, site index in the batch): the file is
    # one batch either way, so the bytes are equal
    assert FEAT["n"] <= 65536 and outs[0].count(b"\n") == FEAT["n"] and outs[0] == outs[1]
    # other batch cuts change the random initial states, not the sites, their order or (beyond the model's own
    # sensitivity to h0/c0) the probabilities
    small = str(tmp_path / "small.dspf")
    feature_bin.pack_feature_file(text, small, batch_sites=300)
    out = str(tmp_path / "small.calls")
    cli.main(["call_mods", "-i", small, "-m", ckpt, "-o", out, "--max_batch", "256"])
    a = [l.split(b"\t") for l in outs[0].splitlines()]
    b = [l.split(b"\t") for l in open(out, "rb").read().splitlines()]
    assert len(a) == len(b) and all(x[:6] == y[:6] and x[9] == y[9] for x, y in zip(a, b))
    assert max(abs(float(x[7]) - float(y[7])) for x, y in zip(a, b)) < 0.05


@pytest.mark.gpu
def test_extract_writes_a_binary_file_equal_to_its_packed_text_file(tmp_path):
    from deepsignal_plant_b200 import cli, extract_features as ef, synthetic
    reads = synthetic.make_reads(12, seed=5, mean_bases=600)
    arc = str(tmp_path / "reads.npz")
    ef.save_reads(arc, reads)
    text, binary, packed = str(tmp_path / "f.tsv"), str(tmp_path / "f.dspf"), str(tmp_path / "p.dspf")
    for out in (text, binary):
        cli.main(["extract", "-i", arc, "-o", out, "--motifs", "CG", "--f5_batch_size", "5"])
    feature_bin.pack_feature_file(text, packed)
    a, ia = collect(feature_bin.FeatureBinReader(binary, pinned=False))
    b, ib = collect(feature_bin.FeatureBinReader(packed, pinned=False))
    assert ia == ib and len(ia) > 0
    for k in a:
        assert a[k].tobytes() == b[k].tobytes(), k


class _HostOnlyModel:
    """Takes the place of ModelBiLSTM for the host-side pipeline: prob_1 = a deterministic function of the batch."""
    num_classes = 2

    def __init__(self):
        self.calls = 0

    @staticmethod
    def answer(means):
        p1 = (np.abs(means.numpy().astype(np.float64)).sum(1) % 1.0).astype(np.float32)
        return np.stack([1 - p1, p1], 1), (p1 > 0.5).astype(np.int32)

    def submit_host(self, kmer, means, stds, lens, signals, logits, probs, labels):
        p, lab = self.answer(means)
        probs.numpy()[:] = p
        labels.numpy()[:] = lab
        self.calls += 1
        return self.calls

    def wait_host(self, ticket):
        return None


def _repack(src, dst, block_sizes):
    """Copy a .dspf file into blocks of the given sizes (cycled)."""
    rd = feature_bin.FeatureBinReader(src, batch_sites=1 << 20, pinned=False, slots=1)
    b = next(iter(rd))
    arrs = [t.numpy() for t in b.arrays()] + [b.labels.numpy()]
    with feature_bin.FeatureBinWriter(dst, rd.T, rd.S) as w:
        a, k = 0, 0
        while a < b.n:
            z = min(a + block_sizes[k % len(block_sizes)], b.n)
            w.write(*[x[a:z] for x in arrs], b.info_text, b.info_off[a:z + 1])
            a, k = z, k + 1


@pytest.mark.parametrize("batch_sites,block_sizes", [(16, (64,)), (100, (64,)), (128, (10, 64, 5, 100, 63))])
def test_call_mods_stream_writes_batches_in_order_while_slots_recycle(tmp_path, batch_sites, block_sizes):
    from deepsignal_plant_b200 import call_modifications as cm
    p0, p = str(tmp_path / "f0.dspf"), str(tmp_path / "f.dspf")
    feature_bin.pack_feature_file(TEXT, p0, nthreads=2)
    _repack(p0, p, block_sizes)                                  # batches of unequal sizes: result slots of several sizes
    want = b"".join(feature_io.format_calls(b, *_HostOnlyModel.answer(b.base_means))
                    for b in feature_bin.FeatureBinReader(p, batch_sites=batch_sites, pinned=False, slots=2))
    got = []
    rd = feature_bin.FeatureBinReader(p, batch_sites=batch_sites, pinned=False, slots=8, nthreads=2)
    sites, acc, nb = cm.call_mods_stream(_HostOnlyModel(), rd, lambda a: got.append(bytes(a)))
    assert sites == FEAT["n"] and nb == len(got) and 0.0 <= acc <= 1.0
    assert b"".join(got) == want and want.count(b"\n") == FEAT["n"]


def test_call_mods_stream_surfaces_a_failing_writer(tmp_path):
    from deepsignal_plant_b200 import call_modifications as cm
    p = str(tmp_path / "f.dspf")
    feature_bin.pack_feature_file(TEXT, p, nthreads=2)

    def broken(_):
        raise OSError("disk full")
    with pytest.raises(OSError, match="disk full"):
        cm.call_mods_stream(_HostOnlyModel(), feature_bin.FeatureBinReader(p, batch_sites=32, pinned=False, slots=8), broken)


@pytest.mark.gpu
def test_two_ranks_take_contiguous_site_ranges_of_a_binary_file(tmp_path):
    # torchrun, 2 ranks (sharing the GPU here): every rank reads its site range of the .dspf file, rank 0 concatenates the
    # parts in rank order -> the sites of the file, in file order, exactly once
    import subprocess
    import sys
    import torch
    from deepsignal_plant_b200.models import ModelBiLSTM
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    binary = str(tmp_path / "f.dspf")
    feature_bin.pack_feature_file(TEXT, binary, batch_sites=100)
    ckpt = str(tmp_path / "m.ckpt")
    torch.manual_seed(1234)
    torch.save(ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True).state_dict(), ckpt)
    out = str(tmp_path / "calls.tsv")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29741", "-m", "deepsignal_plant_b200", "call_mods", "-i", binary, "-m", ckpt, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    _, info = collect(feature_bin.FeatureBinReader(binary, pinned=False))
    lines = open(out).read().splitlines()
    assert len(lines) == FEAT["n"] and ["\t".join(l.split("\t")[:6]) for l in lines] == info
    assert all(abs(float(l.split("\t")[6]) + float(l.split("\t")[7]) - 1.0) < 2e-6 for l in lines)
