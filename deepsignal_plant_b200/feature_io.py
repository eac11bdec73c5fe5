"""Feature-file ingestion and call_mods output formatting around the hot path.

Mirrors, with whole-batch native conversions instead of per-line Python:

* ``_read_features_file`` (``deepsignal_plant/call_modifications.py:55-127``): the 12-column
  feature TSV written by ``deepsignal_plant extract`` (``extract_features.py:381-395``) ->
  batches of the five float32 arrays ``ModelBiLSTM.forward`` takes, parsed by
  ``dsp_parse_features`` straight into page-locked buffers;
* the per-site text loop of ``_call_mods`` (``:175-188``) and ``_write_predstr_to_file``
  (``:262-282``): ``dsp_format_calls`` renders a batch's lines into one bytes object.

The reference groups lines into batches of ``f5_batch_size`` reads and hands Python lists across
process queues; here a batch is ``batch_sites`` consecutive lines (file order is kept end to end).
"""
from __future__ import annotations

import ctypes as C
import gzip
import os

import numpy as np

from . import _native

code2base_dna = "ACGTNWSMKRYBVDHZ"


def features_to_str(sampleinfo, kmer_codes, means, stds, lens, signals, label):
    """One feature-file line, as ``_features_to_str`` writes it (``extract_features.py:381-395``):
    values already rounded to 6 decimals are printed with ``str``."""
    k_mer = "".join(code2base_dna[int(c)] for c in kmer_codes)
    means_text = ",".join(str(x) for x in np.around(np.asarray(means, dtype=np.float64), decimals=6))
    stds_text = ",".join(str(x) for x in np.around(np.asarray(stds, dtype=np.float64), decimals=6))
    len_text = ",".join(str(int(x)) for x in lens)
    sig_text = ";".join(",".join(str(y) for y in np.asarray(x, dtype=np.float64)) for x in signals)
    return "\t".join([sampleinfo, k_mer, means_text, stds_text, len_text, sig_text, str(int(label))])


class FeatureBatch:
    """One parsed block of the feature file.  ``kmer`` ... ``signals`` are float32 torch tensors
    (page-locked when ``pinned``) trimmed to ``n`` sites; ``labels`` int32; ``text`` keeps the
    block's bytes alive for the output lines (sampleinfo columns and the k-mer are copied from it)."""
    __slots__ = ("n", "kmer", "base_means", "base_stds", "base_signal_lens", "signals", "labels",
                 "text", "line_begin", "info_len", "kmer_off", "seq_len", "slot")

    def arrays(self):
        return self.kmer, self.base_means, self.base_stds, self.base_signal_lens, self.signals

    def sampleinfo(self):
        """list of the tab-joined first six columns (``call_modifications.py:89``)."""
        t = self.text
        return [t[b:b + l].decode() for b, l in zip(self.line_begin[:self.n].tolist(), self.info_len[:self.n].tolist())]

    def as_reference_lists(self):
        """The 7-tuple of Python lists ``_read_features_file`` puts on its queue (``:103-104``)."""
        n = self.n
        return (self.sampleinfo(), self.kmer[:n].numpy().astype(np.int64).tolist(), self.base_means[:n].numpy().tolist(),
                self.base_stds[:n].numpy().tolist(), self.base_signal_lens[:n].numpy().astype(np.int64).tolist(),
                self.signals[:n].numpy().tolist(), self.labels[:n].numpy().tolist())


def shard_bounds(f, lo, hi):
    """Byte interval [start, end) of the lines of binary file ``f`` that START in [lo, hi)."""
    size = f.seek(0, os.SEEK_END)

    def first_line_start_at_or_after(x):
        if x <= 0:
            return 0
        if x >= size:
            return size
        f.seek(x - 1)
        return x - 1 + len(f.readline())
    return first_line_start_at_or_after(lo), first_line_start_at_or_after(hi)


class _Slot:
    def __init__(self, cap, T, S, pinned):
        import torch
        mk = lambda shape, dt: (torch.empty(shape, dtype=dt).pin_memory() if pinned else torch.empty(shape, dtype=dt))
        self.kmer, self.means, self.stds, self.lens = (mk((cap, T), torch.float32) for _ in range(4))
        self.signals = mk((cap, T, S), torch.float32)
        self.labels = mk((cap,), torch.int32)
        self.line_begin = np.empty(cap, np.int64)
        self.info_len = np.empty(cap, np.int32)
        self.kmer_off = np.empty(cap, np.int32)


class FeatureFileReader:
    """Iterate ``FeatureBatch`` objects over a feature file (plain or ``.gz``).

    ``slots`` buffers are cycled: a batch's tensors are valid until ``slots - 1`` further batches
    have been produced (or ``release(batch)`` is called when ``blocking_pool`` is used by the
    pipeline).  ``byte_range=(lo, hi)`` restricts a plain file to the lines that START in
    ``[lo, hi)`` -- contiguous shards for one-process-per-GPU runs."""

    def __init__(self, path, seq_len=13, signal_len=16, batch_sites=65536, pinned=None, slots=4,
                 nthreads=None, byte_range=None):
        import torch
        self.path, self.T, self.S = path, int(seq_len), int(signal_len)
        self.batch_sites = int(batch_sites)
        self.pinned = torch.cuda.is_available() if pinned is None else bool(pinned)
        self.nthreads = int(nthreads or min(32, os.cpu_count() or 1))
        self.nslots = int(slots)
        self._slots = [None] * self.nslots
        self.byte_range = byte_range
        if byte_range is not None and path.endswith(".gz"):
            raise ValueError("byte_range needs an uncompressed feature file")
        self.sites_read = 0

    def _slot(self, i):
        if self._slots[i] is None:
            self._slots[i] = _Slot(self.batch_sites, self.T, self.S, self.pinned)
        return self._slots[i]

    def __iter__(self):
        L = _native.lib()
        f = gzip.open(self.path, "rb") if self.path.endswith(".gz") else open(self.path, "rb")
        remaining = None
        with f:
            if self.byte_range is not None:
                start, end = shard_bounds(f, *self.byte_range)
                f.seek(start)
                remaining = end - start
            buf = b""
            eof = False
            est_line = 4 * self.T * 10 + self.T * self.S * 10       # refined after the first block
            k = 0
            while True:
                want = int(est_line * self.batch_sites * 1.05) + (1 << 16)
                while not eof and len(buf) < want:
                    ask = want - len(buf) if remaining is None else min(want - len(buf), remaining)
                    chunk = f.read(ask) if ask > 0 else b""
                    if not chunk:
                        eof = True
                        break
                    if remaining is not None:
                        remaining -= len(chunk)
                    buf += chunk
                if not buf or buf.isspace():
                    return
                s = self._slot(k % self.nslots)
                n, used = C.c_int64(0), C.c_int64(0)
                _native.check(L.dsp_parse_features(
                    buf, len(buf), int(eof), self.T, self.S, self.batch_sites,
                    s.kmer.data_ptr(), s.means.data_ptr(), s.stds.data_ptr(), s.lens.data_ptr(), s.signals.data_ptr(),
                    s.labels.data_ptr(), s.line_begin.ctypes.data, s.info_len.ctypes.data, s.kmer_off.ctypes.data,
                    C.byref(n), C.byref(used), self.nthreads), "dsp_parse_features(%s)" % self.path)
                n, used = int(n.value), int(used.value)
                if n == 0:
                    if eof:
                        return
                    est_line *= 2                        # a line longer than the whole block: read more
                    continue
                b = FeatureBatch()
                b.n, b.seq_len, b.slot = n, self.T, k % self.nslots
                b.kmer, b.base_means, b.base_stds, b.base_signal_lens = s.kmer[:n], s.means[:n], s.stds[:n], s.lens[:n]
                b.signals, b.labels = s.signals[:n], s.labels[:n]
                b.text = buf[:used]
                b.line_begin, b.info_len, b.kmer_off = s.line_begin, s.info_len, s.kmer_off
                buf = buf[used:]
                est_line = max(64, used // n)
                self.sites_read += n
                k += 1
                yield b


def format_calls(batch, probs, labels, nthreads=None):
    """bytes of the call_mods lines of one batch (``call_modifications.py:175-188``), each
    terminated by a newline (what ``_write_predstr_to_file`` writes, ``:279-280``).
    ``probs`` (n, 2) float32 and ``labels`` (n) int32: numpy arrays or CPU torch tensors."""
    L = _native.lib()
    n = batch.n
    p = np.ascontiguousarray(np.asarray(probs, dtype=np.float32))
    lab = np.ascontiguousarray(np.asarray(labels, dtype=np.int32))
    if p.shape != (n, 2) or lab.shape != (n,):
        raise ValueError("format_calls: probs must be (n, 2) and labels (n,) for the batch's n = %d" % n)
    cap = int(batch.info_len[:n].sum()) + n * (64 + batch.seq_len)
    out = C.create_string_buffer(cap)
    used = C.c_int64(0)
    _native.check(L.dsp_format_calls(batch.text, batch.line_begin.ctypes.data, batch.info_len.ctypes.data,
                                     batch.kmer_off.ctypes.data, batch.seq_len, p.ctypes.data, lab.ctypes.data, n,
                                     out, cap, C.byref(used), int(nthreads or min(16, os.cpu_count() or 1))),
                  "dsp_format_calls")
    return out.raw[:used.value]
