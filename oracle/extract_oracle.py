"""ORACLE (test infrastructure, not product code): restatement of the reference's feature
extraction from a re-squiggled read, starting at the decoded arrays (raw DAC samples, channel
scaling, tombo event table) -- SURVEY.md section 8(f) row 4.

Follows ``deepsignal_plant/extract_features.py``:
  * ``rescale_signals``      <- ``_rescale_signals`` (:276-277)
  * ``normalize_signals``    <- ``_normalize_signals`` (:179-190)
  * ``signals_rect``         <- ``_get_signals_rect`` (:232-251)
  * ``extract_features``     <- the per-read body of ``_extract_features`` (:280-378), with the three
                                fast5 accessors (``_get_alignment_info_from_fast5`` :150-176,
                                ``_get_label_raw`` :37-91, ``_get_scaling_of_a_read`` :255-273) replaced
                                by fields of an already decoded read (h5py is absent here)
  * ``find_sites``           <- ``get_refloc_of_methysite_in_motif`` (utils/process_utils.py:97-112)
  * ``get_motif_seqs``       <- ``get_motif_seqs`` / ``_convert_motif_seq`` (utils/process_utils.py:115-147)

Third-party arithmetic the reference calls and this file calls too (numpy): ``np.median``, ``np.mean``
and ``np.std`` (float64, numpy's pairwise summation), ``np.around(x, 6)`` (= rint(x*1e6)/1e6).
Third-party arithmetic that is ABSENT here and restated: ``statsmodels.robust.mad`` (statsmodels is
pinned ``>=0.9.0`` in the reference's requirements.txt, not installed in this image).  Its published
definition, ``mad(a, c=scipy.stats.norm.ppf(3/4.), axis=0, center=np.median)`` =
``np.median(np.abs(a - center(a)) / c)``, is restated in ``mad`` below.  Pin status: everything but that
one function is pinned by ``tests/golden/extract_*.npz``, which ``oracle/make_golden_extract.py``
produces by running the reference's unmodified ``_extract_features`` on synthetic decoded reads (its
fast5 accessors monkeypatched to serve them, ``statsmodels.robust.mad`` bound to the restatement);
the MAD constant itself is "parity unpinned" (restated from the published algorithm).

A decoded read is a dict: readname, strand ('t'/'c'), alignstrand ('+'/'-'), chrom, chrom_start,
raw (int16 array), scaling / offset (float64 or None), ev_start (int array, already shifted by
``read_start_rel_to_raw`` as :80 does), ev_len (int array), ev_base (str, one letter per event).

The ordered random subsample of a base with more than ``signals_len`` samples (:247-249) draws from
Python's global ``random``; ``extract_features`` takes the generator as an argument and also returns
the offsets it drew, so that a caller can replay them (the CUDA path's parity mode).
"""
from __future__ import annotations

import random as _random

import numpy as np

MAD_C = 0.6744897501960817      # scipy.stats.norm.ppf(3/4.), the default `c` of statsmodels.robust.mad

iupac_alphabets = {'A': ['A'], 'T': ['T'], 'C': ['C'], 'G': ['G'], 'R': ['A', 'G'], 'M': ['A', 'C'],
                   'S': ['C', 'G'], 'Y': ['C', 'T'], 'K': ['G', 'T'], 'W': ['A', 'T'], 'B': ['C', 'G', 'T'],
                   'D': ['A', 'G', 'T'], 'H': ['A', 'C', 'T'], 'V': ['A', 'C', 'G'], 'N': ['A', 'C', 'G', 'T']}
base2code_dna = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'N': 4, 'W': 5, 'S': 6, 'M': 7, 'K': 8, 'R': 9,
                 'Y': 10, 'B': 11, 'V': 12, 'D': 13, 'H': 14, 'Z': 15}


def get_motif_seqs(motifs):
    out = []
    for ori in motifs.strip().split(','):
        seqs = ['']
        for b in ori.strip().upper():
            seqs = [s + x for s in seqs for x in iupac_alphabets[b]]
        out += seqs
    return out


def find_sites(seqstr, motifset, methyloc_in_motif=0):
    motifset = set(motifset)
    mlen = len(next(iter(motifset)))
    return [i + methyloc_in_motif for i in range(0, len(seqstr) - mlen + 1) if seqstr[i:i + mlen] in motifset]


def rescale_signals(raw, scaling, offset):
    return np.array(scaling * (raw + offset), dtype=float)


def mad(a):
    a = np.asarray(a)
    center = np.median(a)
    return np.median(np.abs(a - center) / MAD_C)


def normalize_signals(signals, method="mad"):
    if method == "zscore":
        sshift, sscale = np.mean(signals), float(np.std(signals))
    elif method == "mad":
        sshift, sscale = np.median(signals), float(mad(signals))
    else:
        raise ValueError("")
    norm = signals if sscale == 0.0 else (signals - sshift) / sscale
    return np.around(norm, decimals=6)


def signals_rect(signals_list, signals_len=16, rng=_random):
    """-> (rectangle rows, offsets drawn per row or None)."""
    rect, drawn = [], []
    for tmp in signals_list:
        signals = list(np.around(tmp, decimals=6))
        idx = None
        if len(signals) < signals_len:
            pad = signals_len - len(signals)
            left = pad // 2
            signals = [0.] * left + signals + [0.] * (pad - left)
        elif len(signals) > signals_len:
            idx = sorted(rng.sample(range(len(signals)), signals_len))
            signals = [signals[x] for x in idx]
        rect.append(signals)
        drawn.append(idx)
    return rect, drawn


def extract_features(reads, normalize_method, motif_seqs, methyloc, chrom2len, kmer_len, signals_len,
                     methy_label, positions=None, regioninfo=(None, None, None), rng=_random):
    """-> (features_list as the reference builds it, drawn) where drawn[i][j] is the sorted offset list
    the j-th base of site i was subsampled with (None when it had <= signals_len samples)."""
    if kmer_len % 2 == 0:
        raise ValueError("kmer_len must be odd")
    num_bases = (kmer_len - 1) // 2
    features_list, drawn_all = [], []
    rg_chrom, rg_start, rg_end = regioninfo
    for rd in reads:
        chrom, chrom_start, alignstrand = rd["chrom"], rd["chrom_start"], rd["alignstrand"]
        if rg_chrom is not None and rg_chrom != chrom:
            continue
        raw = rd["raw"]
        if rd.get("scaling") is not None:
            raw = rescale_signals(raw, rd["scaling"], rd["offset"])
        norm = normalize_signals(raw, normalize_method)
        genomeseq = "".join(rd["ev_base"])
        signal_list = [norm[s:s + l] for s, l in zip(rd["ev_start"], rd["ev_len"])]
        read_rg_start = chrom_start if rg_start is None else rg_start
        read_rg_end = chrom_start + len(genomeseq) if rg_end is None else rg_end
        if read_rg_start >= chrom_start + len(genomeseq) or read_rg_end <= chrom_start:
            continue
        chromlen = None
        if chrom2len is not None:
            chromlen = chrom2len.get(chrom)
        for loc in find_sites(genomeseq, set(motif_seqs), methyloc):
            if not (num_bases <= loc < len(genomeseq) - num_bases):
                continue
            if alignstrand == '-':
                pos = chrom_start + len(genomeseq) - 1 - loc
                pos_in_strand = chromlen - 1 - pos if chromlen is not None else -1
            else:
                pos = chrom_start + loc
                pos_in_strand = pos if chromlen is not None else -1
            if (rg_chrom is not None) and (pos < read_rg_start or pos >= read_rg_end):
                continue
            if (positions is not None) and ("||".join([chrom, str(pos), alignstrand]) not in positions):
                continue
            k_mer = genomeseq[loc - num_bases:loc + num_bases + 1]
            k_signals = signal_list[loc - num_bases:loc + num_bases + 1]
            lens = [len(x) for x in k_signals]
            means = [np.mean(x) for x in k_signals]
            stds = [np.std(x) for x in k_signals]
            rect, drawn = signals_rect(k_signals, signals_len, rng)
            features_list.append((chrom, pos, alignstrand, pos_in_strand, rd["readname"], rd["strand"],
                                  k_mer, means, stds, lens, rect, methy_label))
            drawn_all.append(drawn)
    return features_list, drawn_all


def features_to_arrays(features_list, round_stats):
    """The five model inputs as ``FloatTensor`` would see them (float32).  round_stats=True is the
    feature-FILE route (``_features_to_str`` rounds means/stds to 6 decimals, :388-389, the reader
    parses them back); False is the direct fast5 route (``_read_features_from_fast5s``,
    call_modifications.py:309-318: unrounded float64 -> float32)."""
    n = len(features_list)
    if n == 0:
        return None
    T = len(features_list[0][6])
    S = len(features_list[0][10][0])
    kmer = np.zeros((n, T), np.float32)
    means = np.zeros((n, T), np.float32)
    stds = np.zeros((n, T), np.float32)
    lens = np.zeros((n, T), np.float32)
    sig = np.zeros((n, T, S), np.float32)
    for i, f in enumerate(features_list):
        kmer[i] = [base2code_dna[c] for c in f[6]]
        m, s = np.asarray(f[7], np.float64), np.asarray(f[8], np.float64)
        if round_stats:
            m, s = np.around(m, 6), np.around(s, 6)
        means[i], stds[i], lens[i] = m, s, f[9]
        sig[i] = np.asarray(f[10], np.float64)
    return dict(kmer=kmer, base_means=means, base_stds=stds, base_signal_lens=lens, signals=sig)


def drawn_to_array(drawn_all, T, S):
    """(n, T, S) int32 of subsample offsets, -1 rows where nothing was drawn."""
    out = np.full((len(drawn_all), T, S), -1, np.int32)
    for i, d in enumerate(drawn_all):
        for j, idx in enumerate(d):
            if idx is not None:
                out[i, j] = idx
    return out


def pack_reads(reads):
    """Decoded reads -> flat arrays (the layout the product API takes)."""
    raw_off = np.concatenate([[0], np.cumsum([len(r["raw"]) for r in reads])]).astype(np.int64)
    ev_off = np.concatenate([[0], np.cumsum([len(r["ev_len"]) for r in reads])]).astype(np.int64)
    return dict(
        raw=np.concatenate([r["raw"] for r in reads]).astype(np.int16), raw_off=raw_off, ev_off=ev_off,
        scaling=np.array([np.nan if r["scaling"] is None else r["scaling"] for r in reads], np.float64),
        offset=np.array([0.0 if r["offset"] is None else r["offset"] for r in reads], np.float64),
        ev_start=np.concatenate([r["ev_start"] for r in reads]).astype(np.int64),
        ev_len=np.concatenate([r["ev_len"] for r in reads]).astype(np.int64),
        ev_base=np.frombuffer("".join(r["ev_base"] for r in reads).encode(), np.uint8).copy(),
        readname=np.array([r["readname"] for r in reads]), strand=np.array([r["strand"] for r in reads]),
        alignstrand=np.array([r["alignstrand"] for r in reads]), chrom=np.array([r["chrom"] for r in reads]),
        chrom_start=np.array([r["chrom_start"] for r in reads], np.int64))


def unpack_reads(z):
    """Inverse of pack_reads (tests rebuild the dict form from a fixture with this)."""
    reads = []
    for i in range(len(z["readname"])):
        a, b = int(z["raw_off"][i]), int(z["raw_off"][i + 1])
        c, d = int(z["ev_off"][i]), int(z["ev_off"][i + 1])
        sc = z["scaling"][i]
        reads.append(dict(readname=str(z["readname"][i]), strand=str(z["strand"][i]), alignstrand=str(z["alignstrand"][i]),
                          chrom=str(z["chrom"][i]), chrom_start=int(z["chrom_start"][i]), raw=z["raw"][a:b],
                          scaling=None if np.isnan(sc) else np.float64(sc),
                          offset=None if np.isnan(sc) else np.float64(z["offset"][i]),
                          ev_start=z["ev_start"][c:d], ev_len=z["ev_len"][c:d],
                          ev_base=bytes(z["ev_base"][c:d]).decode()))
    return reads
