"""Synthetic per-site features of the deepsignal-plant shape.

There is no network and no fast5 data here, so every test and benchmark runs on
features drawn from the distributions SURVEY.md fixes:

* ``kmer``: ``seq_len`` base codes uniform in {0,1,2,3} (A,C,G,T of
  ``base2code_dna``, reference ``utils/process_utils.py:22-25``) with the centre
  forced to 1 ('C'); float32-coded, as ``call_modifications.py:159`` passes it.
* ``base_means`` ~ N(0,1), ``base_stds`` ~ |N(0.3,0.1)|, rounded to 6 decimals
  (``extract_features.py:386-387``).
* ``base_signal_lens``: integers uniform in [3,40), float32-coded.
* ``signals`` (seq_len, signal_len): per base ``len`` samples ~ N(mean,std), then
  the reference's rectangle rule (``extract_features.py:232-251``): centred zero
  padding if ``len < signal_len``, ordered random subsample if ``len > signal_len``.

numpy's ``default_rng`` streams are stable across platforms, so the same seed
gives the same bytes here and on the GPU box.
"""
from __future__ import annotations

import numpy as np

MAX_BASE_LEN = 40  # exclusive upper bound of base_signal_lens


def _rectangle(raw, lens, signal_len, rng):
    """Vectorised restatement of the rectangle rule for a (n, T, L) block.

    raw[..., j] is valid for j < lens; returns (n, T, signal_len).
    """
    n, T, L = raw.shape
    idx = np.arange(L)[None, None, :]
    valid = idx < lens[..., None]
    # --- subsample branch: keep `signal_len` of the `len` samples, original order
    keys = rng.random((n, T, L))
    keys[~valid] = np.inf
    # rank of each key among the valid ones
    order = np.argsort(keys, axis=-1, kind="stable")
    rank = np.empty_like(order)
    np.put_along_axis(rank, order, np.broadcast_to(idx, order.shape).copy(), axis=-1)
    keep = valid & (rank < signal_len)
    # stable compaction: positions of kept samples, in order, moved to the front
    front = np.argsort(~keep, axis=-1, kind="stable")[..., :signal_len]
    sub = np.take_along_axis(raw, front, axis=-1)
    nkeep = np.minimum(lens, signal_len)
    sub[np.arange(signal_len)[None, None, :] >= nkeep[..., None]] = 0.0
    # --- pad branch: shift right by pad_left = (signal_len - len) // 2
    pad_left = np.maximum(signal_len - lens, 0) // 2
    out = np.zeros((n, T, signal_len), dtype=raw.dtype)
    src = np.arange(signal_len)[None, None, :] - pad_left[..., None]
    ok = (src >= 0) & (src < nkeep[..., None])
    gathered = np.take_along_axis(sub, np.clip(src, 0, signal_len - 1), axis=-1)
    out[ok] = gathered[ok]
    return out


def make_features(n, seq_len=13, signal_len=16, seed=0, chunk=8192):
    """Return dict of float32 arrays: kmer, base_means, base_stds, base_signal_lens
    (each (n, seq_len)) and signals (n, seq_len, signal_len)."""
    rng = np.random.default_rng(seed)
    T = seq_len
    kmer = np.empty((n, T), np.float32)
    means = np.empty((n, T), np.float32)
    stds = np.empty((n, T), np.float32)
    lens = np.empty((n, T), np.float32)
    signals = np.empty((n, T, signal_len), np.float32)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        k = rng.integers(0, 4, (m, T))
        k[:, T // 2] = 1
        mu = np.round(rng.standard_normal((m, T)), 6)
        sd = np.round(np.abs(rng.normal(0.3, 0.1, (m, T))), 6)
        ln = rng.integers(3, MAX_BASE_LEN, (m, T))
        raw = mu[..., None] + sd[..., None] * rng.standard_normal((m, T, MAX_BASE_LEN - 1))
        raw = np.round(raw, 6)
        rect = _rectangle(raw, ln, signal_len, rng)
        kmer[s:s + m] = k
        means[s:s + m] = mu
        stds[s:s + m] = sd
        lens[s:s + m] = ln
        signals[s:s + m] = rect
    return {"kmer": kmer, "base_means": means, "base_stds": stds,
            "base_signal_lens": lens, "signals": signals}


def make_sampleinfo(n, seed=0, n_chrom=5, n_pos=100000):
    """Six leading call_mods columns per site (``call_modifications.py:85``):
    chrom, pos, strand, pos_in_strand, readname, read_strand joined by tabs."""
    rng = np.random.default_rng(seed + 7919)
    chrom = rng.integers(1, n_chrom + 1, n)
    pos = rng.integers(0, n_pos, n)
    read = rng.integers(0, max(1, n // 20), n)
    out = []
    for c, p, r in zip(chrom.tolist(), pos.tolist(), read.tolist()):
        strand = "+" if (p & 1) == 0 else "-"
        pis = p if strand == "+" else n_pos - 1 - p
        out.append("chr%d\t%d\t%s\t%d\tread_%06d\tt" % (c, p, strand, pis, r))
    return out


def make_states(cfg, n, seed=4321):
    """Explicit initial LSTM states for parity runs, drawn in the order the reference
    calls ``init_hidden`` (``models.py:196,212,226``): seq h0,c0; signal h0,c0; comb h0,c0.
    Returned as float32 numpy arrays keyed 'seq'/'signal'/'comb' -> (h0, c0), each
    (num_layers*2, n, hidden) indexed [layer*2 + dir] like ``nn.LSTM``."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = {}
    if cfg["module"] != "signal_bilstm":
        out["seq"] = tuple(torch.randn(cfg["num_layers2"] * 2, n, cfg["nhid_seq"], generator=g).numpy()
                           for _ in range(2))
    if cfg["module"] != "seq_bilstm":
        out["signal"] = tuple(torch.randn(cfg["num_layers2"] * 2, n, cfg["nhid_signal"], generator=g).numpy()
                              for _ in range(2))
    out["comb"] = tuple(torch.randn(cfg["num_layers1"] * 2, n, cfg["hidden_size"], generator=g).numpy()
                        for _ in range(2))
    return out


def make_callmods_records(n_records, n_chrom=5, n_pos=10000, seed=0, tie_fraction=0.02):
    """Synthetic per-read call_mods lines (10 columns, ``call_modifications.py:1-6``)
    for the call_freq path: strand = pos & 1, probabilities printed the way
    ``_call_mods`` prints them (float32 rounded to 6 dp, shortest repr), with a
    fraction of near-0.5 calls so that ``--prob_cf`` filters something."""
    rng = np.random.default_rng(seed)
    chrom = rng.integers(1, n_chrom + 1, n_records)
    pos = rng.integers(0, n_pos, n_records)
    p1 = rng.beta(0.5, 0.5, n_records).astype(np.float32)
    tie = rng.random(n_records) < tie_fraction
    p1[tie] = np.float32(0.5) + (rng.integers(-3, 4, int(tie.sum())) * np.float32(1e-6)).astype(np.float32)
    # a few extreme probabilities that print in scientific notation
    ext = rng.random(n_records) < 0.01
    p1[ext] = (rng.integers(0, 99, int(ext.sum())) * np.float32(1e-6)).astype(np.float32)
    p0n = np.rint((np.float32(1.0) - p1) * np.float32(1e6)) / np.float32(1e6)
    p0n = p0n.astype(np.float32)
    p1n = (np.rint((np.float32(1.0) - p0n) * np.float32(1e6)) / np.float32(1e6)).astype(np.float32)
    label = (p1 > 0.5).astype(np.int64)
    bases = "ACGT"
    kmers = rng.integers(0, 4, (n_records, 5))
    kmers[:, 2] = 1
    lines = []
    for i in range(n_records):
        p = int(pos[i])
        strand = "+" if (p & 1) == 0 else "-"
        pis = p if strand == "+" else n_pos - 1 - p
        lines.append("chr%d\t%d\t%s\t%d\tread_%07d\tt\t%s\t%s\t%d\t%s" % (
            chrom[i], p, strand, pis, i // 25, str(p0n[i]), str(p1n[i]), label[i],
            "".join(bases[b] for b in kmers[i])))
    return lines


def make_reads(n_reads, seed=0, mean_bases=400, mean_dwell=9.0, n_chrom=3, chrom_len=200000, long_every=7,
               no_scaling_every=0, stall_every=0):
    """Synthetic re-squiggled reads, decoded the way ``_get_label_raw`` / ``_get_scaling_of_a_read``
    (``extract_features.py:37-91,255-273``) hand them to ``_extract_features``: int16 DAC samples, the
    channel's scaling / offset, and the tombo event table (start already shifted by
    ``read_start_rel_to_raw``, length, base).  Dwell times are geometric-like with a heavy tail so that
    bases shorter than, equal to and longer than the 16-sample rectangle all occur; every
    ``long_every``-th read has a stalled stretch (dwell > 100 samples: numpy's pairwise summation
    switches to its 8-accumulator and recursive forms there); ``stall_every`` adds a base of 900-2600 samples.
    Returns a list of dicts."""
    rng = np.random.default_rng(seed)
    levels = rng.normal(0.0, 1.0, 4 ** 3)                       # a 3-mer pore model in normalised units
    reads = []
    for r in range(n_reads):
        nb = int(max(30, rng.poisson(mean_bases)))
        seq = rng.integers(0, 4, nb)
        dwell = 1 + rng.geometric(1.0 / mean_dwell, nb) - 1
        dwell = np.maximum(dwell, 1)
        tail = rng.random(nb) < 0.04
        dwell[tail] += rng.integers(8, 60, int(tail.sum()))
        if long_every and r % long_every == 0:
            j = rng.integers(0, nb, 3)
            dwell[j] = rng.integers(110, 400, 3)
        if stall_every and r % stall_every == 0:                 # a blocked pore: one base dwells for thousands of samples
            dwell[int(rng.integers(10, nb - 10))] = int(rng.integers(900, 2600))
        lead = int(rng.integers(0, 300))                         # read_start_rel_to_raw
        trail = int(rng.integers(0, 200))
        starts = lead + np.concatenate([[0], np.cumsum(dwell)[:-1]])
        total = lead + int(dwell.sum()) + trail
        ctx = np.zeros(nb, np.int64)
        ctx[1:-1] = seq[:-2] * 16 + seq[1:-1] * 4 + seq[2:]
        per_base = levels[ctx]
        norm = np.repeat(per_base, dwell) + rng.normal(0, 0.25, int(dwell.sum()))
        full = np.concatenate([rng.normal(0.5, 1.0, lead), norm, rng.normal(-0.3, 1.2, trail)])
        scaling = float(rng.uniform(0.15, 0.2))
        offset = float(rng.integers(-20, 30))
        pa = 90.0 + 12.0 * full                                  # picoampere
        raw = np.clip(np.rint(pa / scaling - offset), -32768, 32767).astype(np.int16)
        assert raw.shape[0] == total
        rd = dict(readname="read_%05d" % r, strand="t", alignstrand="+-"[int(rng.integers(0, 2))],
                  chrom="chr%d" % int(rng.integers(1, n_chrom + 1)),
                  chrom_start=int(rng.integers(0, chrom_len - nb)),
                  raw=raw, scaling=np.float64(scaling), offset=np.float64(offset),
                  ev_start=starts.astype(np.int64), ev_len=dwell.astype(np.int64),
                  ev_base="".join("ACGT"[b] for b in seq))
        if no_scaling_every and r % no_scaling_every == 1:
            rd["scaling"], rd["offset"] = None, None
        reads.append(rd)
    return reads
