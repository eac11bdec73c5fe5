"""Multi-rank call_freq (freq_dist.py + csrc/comm.cu).

CPU (``-m "not gpu"``): the host logic under world-size 2 / 3 ``gloo`` with the stand-in backend of
tests/dist_standin.py -- the file written by the ranks must equal the bytes of the reference restatement
(oracle/freq_oracle.py, pinned by the reference's own outputs in tests/golden) for every order mode.
GPU (``-m gpu``): the same command line with the real backend -- fused partition + exchange kernels writing
into CUDA-IPC windows -- run as 2 and 3 ranks that SHARE cuda:0 (IPC works between processes on one device,
the control plane then runs on gloo), so the exchange is exercised on a one-GPU box too; plus the tensor-level
entry against the single-GPU table, bit for bit."""
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases
from deepsignal_plant_b200 import call_mods_freq as cf
from deepsignal_plant_b200 import freq_dist as fd
from deepsignal_plant_b200 import synthetic
from oracle import freq_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "freq_dist_worker.py")


def _run_ranks(world, args, timeout=600, env_extra=None):
    port = 29600 + (os.getpid() * 7 + world * 13 + len(args)) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER] + [str(a) for a in args]
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-5000:]
    return r.stdout


def _write_inputs(tmp_path, lines, pieces=3, gz_last=True):
    """the records as several input files in argument order (the last one gzipped)"""
    cut = [len(lines) * i // pieces for i in range(pieces + 1)]
    files = []
    for i in range(pieces):
        body = "".join(l + "\n" for l in lines[cut[i]:cut[i + 1]])
        if gz_last and i == pieces - 1:
            p = str(tmp_path / ("calls_%d.tsv.gz" % i))
            with gzip.open(p, "wt") as f:
                f.write(body)
        else:
            p = str(tmp_path / ("calls_%d.tsv" % i))
            with open(p, "w") as f:
                f.write(body)
        files.append(p)
    return files


def test_plan_units_covers_every_byte_once_in_order(tmp_path):
    lines = synthetic.make_callmods_records(5000, n_chrom=3, n_pos=100, seed=1)
    files = _write_inputs(tmp_path, lines, pieces=4)
    for world in (1, 2, 3, 8):
        shards = fd.plan_units(files, world)
        got = []
        for units in shards:
            rec = fd.read_units(units)
            got.append(rec)
        rec = cf.Records.concat(got)
        want = cf.parse_lines(lines)
        assert len(rec) == len(want)
        for fld in ("chrom", "pos", "strand", "pos_in_strand", "p0", "p1", "label", "kmer"):
            assert (getattr(rec, fld) == getattr(want, fld)).all(), (world, fld)


def test_shard_view_assigns_each_line_to_the_shard_it_starts_in():
    text = b"".join(b"line%03d\tx\n" % i for i in range(200))
    buf = np.frombuffer(text, np.uint8)
    for world in (2, 3, 7, 64, 500):
        parts = [cf._shard_view(buf, (len(text) * r // world, len(text) * (r + 1) // world)).tobytes() for r in range(world)]
        assert b"".join(parts) == text
        assert all(p == b"" or p.endswith(b"\n") for p in parts)


MODES = [("unsorted", []), ("sorted", ["--sort"]), ("bed_sorted", ["--sort", "--bed"]), ("gz", ["--gzip"]),
         ("contigs", ["--contigs", "chr3,chr10,chr1,chrNone"]), ("contigs_sorted", ["--contigs", "chr3,chr10,chr1", "--sort"])]


def _expected(lines, prob_cf, mode_args):
    is_sort, is_bed = "--sort" in mode_args, "--bed" in mode_args
    if "--contigs" in mode_args:
        contigs = mode_args[mode_args.index("--contigs") + 1].split(",")
        return freq_oracle.render_by_contig(lines, contigs, prob_cf, is_sort, is_bed)
    return freq_oracle.render(freq_oracle.aggregate(lines, prob_cf), is_sort, is_bed)


def _check_mode(tmp_path, world, backend, mode, mode_args, prob_cf, lines):
    files = _write_inputs(tmp_path, lines)
    out = str(tmp_path / ("freq_%s_%d.txt" % (mode, world)))
    _run_ranks(world, ["--backend", backend, "--prob_cf", prob_cf, "-o", out] + mode_args + ["-i"] + files)
    path = out + ".gz" if "--gzip" in mode_args else out
    got = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    want = _expected(lines, prob_cf, mode_args)
    assert got == want, "%s world %d: %d vs %d bytes" % (mode, world, len(got), len(want))
    return got


@pytest.mark.parametrize("world,mode,mode_args", [(2, m, a) for m, a in MODES if m != "bed_sorted"] + [(3, "bed_sorted", ["--sort", "--bed"])])
def test_host_logic_under_gloo_with_standin_backend(tmp_path, world, mode, mode_args):
    lines = synthetic.make_callmods_records(6000, n_chrom=12, n_pos=120, seed=9)
    _check_mode(tmp_path, world, "standin", mode, mode_args, 0.1, lines)


def test_host_logic_reproduces_the_reference_fixture_bytes(tmp_path):
    # the edge-case input of the golden set (ties of %.3f, scientific notation, duplicate keys across strands)
    lines = open(cases.GOLD + "/freq_edge_input.tsv").read().splitlines()
    for name in ("freq_edge_cf0p0_unsorted_tsv", "freq_edge_cf0p0_sorted_tsv", "freq_edge_cf0p5_unsorted_bed"):
        e = cases.MANIFEST["freq"][name]
        args = (["--sort"] if e["sort"] else []) + (["--bed"] if e["bed"] else [])
        files = _write_inputs(tmp_path, lines, pieces=2, gz_last=False)
        out = str(tmp_path / (name + ".txt"))
        _run_ranks(2, ["--backend", "standin", "--prob_cf", e["prob_cf"], "-o", out] + args + ["-i"] + files)
        assert open(out).read() == cases.read_gz(name + ".txt.gz"), name


# ---- GPU: the real backend ------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("world,mode,mode_args", [(2, m, a) for m, a in MODES] + [(3, "unsorted", []), (3, "contigs", MODES[4][1])])
def test_ranks_sharing_one_gpu_write_the_reference_bytes(tmp_path, world, mode, mode_args):
    lines = synthetic.make_callmods_records(40000, n_chrom=12, n_pos=700, seed=10)
    _check_mode(tmp_path, world, "device", mode, mode_args, 0.2, lines)


@pytest.mark.gpu
def test_reference_fixture_bytes_through_the_exchange(tmp_path):
    lines = synthetic.make_callmods_records(100000, n_chrom=12, n_pos=900, seed=5)      # the golden "synth" input
    for name in ("freq_synth_cf0p0_unsorted_tsv", "freq_synth_cf0p5_unsorted_bed", "freq_synth_cf0p0_sorted_tsv"):
        e = cases.MANIFEST["freq"][name]
        args = (["--sort"] if e["sort"] else []) + (["--bed"] if e["bed"] else [])
        files = _write_inputs(tmp_path, lines, pieces=2, gz_last=False)
        out = str(tmp_path / (name + ".txt"))
        _run_ranks(2, ["--backend", "device", "--prob_cf", e["prob_cf"], "-o", out] + args + ["-i"] + files)
        assert open(out).read() == cases.read_gz(name + ".txt.gz"), name


@pytest.mark.gpu
@pytest.mark.parametrize("world,home", [(2, False), (4, False), (3, True)])
def test_tensor_entry_equals_single_gpu_table_bit_for_bit(world, home):
    out = _run_ranks(world, ["--tensor_check", "--records", 3000000, "--prob_cf", 0.3] + (["--home_rows"] if home else []))
    line = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    assert line["ok"] and line["world"] == world and line["sites"] > 10000, line
    sizes = line["rows_per_rank"]
    if home:      # records in random order: nearly every site is first seen in rank 0's shard
        assert sizes[0] > 0.9 * sum(sizes), sizes
    else:         # sampled splitters: every rank holds about 1/world of the table
        assert max(sizes) < 1.25 * sum(sizes) / world and min(sizes) > 0.75 * sum(sizes) / world, sizes


@pytest.mark.gpu
def test_window_overflow_fails_on_every_rank_alike():
    out = _run_ranks(2, ["--tensor_check", "--records", 400000, "--prob_cf", 0.0, "--window_records", 1000, "--expect_overflow"])
    line = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    assert line["ok"] and line["overflow_ranks"] == 2, line


@pytest.mark.gpu
def test_call_freq_command_line_under_torchrun(tmp_path):
    # python -m deepsignal_plant_b200 call_freq under torchrun: WORLD_SIZE > 1 selects the distributed path
    lines = synthetic.make_callmods_records(30000, n_chrom=5, n_pos=500, seed=11)
    files = _write_inputs(tmp_path, lines, pieces=2, gz_last=False)
    out = str(tmp_path / "freq.tsv")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", "-m", "deepsignal_plant_b200", "call_freq", "-i", files[0], "-i", files[1], "-o", out, "--sort", "--prob_cf", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert open(out).read() == freq_oracle.render(freq_oracle.aggregate(lines, 0.0), True, False)
