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
    (page-locked when ``pinned``) trimmed to ``n`` sites; ``labels`` int32; ``info_text`` /
    ``info_off`` hold the first six columns of every line packed back to back (the output lines
    start with them).  Everything lives in a reader slot that is recycled after ``slots - 1``
    further batches."""
    __slots__ = ("n", "kmer", "base_means", "base_stds", "base_signal_lens", "signals", "labels",
                 "info_text", "info_off", "seq_len", "slot")

    def arrays(self):
        return self.kmer, self.base_means, self.base_stds, self.base_signal_lens, self.signals

    def sampleinfo(self):
        """list of the tab-joined first six columns (``call_modifications.py:89``)."""
        off = self.info_off[:self.n + 1].tolist()
        t = self.info_text[:off[-1]].tobytes()
        return [t[a:b].decode() for a, b in zip(off[:-1], off[1:])]

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


def sniff_shape(path):
    """(seq_len, signal_len) of a feature file from its first line: k-mer length and samples per base."""
    with (gzip.open(path, "rt") if path.endswith(".gz") else open(path, "r")) as f:
        words = f.readline().strip().split("\t")
    if len(words) < 12:
        raise ValueError("%s does not look like a feature file (12 tab-separated columns)" % path)
    return len(words[6]), len(words[10].split(";")[0].split(","))


class _Slot:
    def __init__(self, cap, T, S, pinned):
        import torch
        mk = lambda shape, dt: (torch.empty(shape, dtype=dt).pin_memory() if pinned else torch.empty(shape, dtype=dt))
        self.kmer, self.means, self.stds, self.lens = (mk((cap, T), torch.float32) for _ in range(4))
        self.signals = mk((cap, T, S), torch.float32)
        self.labels = mk((cap,), torch.int32)
        self.info_off = np.empty(cap + 1, np.int64)
        self.info_text = np.empty(cap * 96, np.uint8)         # grown on demand (DSP_ERR_NOMEM)


class FeatureFileReader:
    """Iterate ``FeatureBatch`` objects over a feature file (plain or ``.gz``).

    ``slots`` buffers are cycled: a batch's tensors are valid until ``slots - 1`` further batches
    have been produced (or ``release(batch)`` is called when ``blocking_pool`` is used by the
    pipeline).  ``byte_range=(lo, hi)`` restricts a plain file to the lines that START in
    ``[lo, hi)`` -- contiguous shards for one-process-per-GPU runs."""

    def __init__(self, path, seq_len=13, signal_len=16, batch_sites=65536, pinned=None, slots=4,
                 nthreads=None, byte_range=None):
        import torch
        if seq_len is None or signal_len is None:            # take the shape from the first line, as the reference does
            seq_len, signal_len = sniff_shape(path)
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

    def _parse(self, L, s, ptr, nbytes, is_final):
        """One ``dsp_parse_features`` call into slot ``s`` -> (sites, bytes consumed)."""
        n, used = C.c_int64(0), C.c_int64(0)
        while True:
            rc = L.dsp_parse_features(
                ptr, nbytes, int(is_final), self.T, self.S, self.batch_sites,
                s.kmer.data_ptr(), s.means.data_ptr(), s.stds.data_ptr(), s.lens.data_ptr(), s.signals.data_ptr(),
                s.labels.data_ptr(), s.info_text.ctypes.data, s.info_text.size, s.info_off.ctypes.data,
                C.byref(n), C.byref(used), self.nthreads)
            if rc == 4 and s.info_text.size < nbytes:          # DSP_ERR_NOMEM: unusually long sample-info columns
                s.info_text = np.empty(min(nbytes, s.info_text.size * 4), np.uint8)
                continue
            _native.check(rc, "dsp_parse_features(%s)" % self.path)
            return int(n.value), int(used.value)

    def _batch(self, s, n, k):
        b = FeatureBatch()
        b.n, b.seq_len, b.slot = n, self.T, k % self.nslots
        b.kmer, b.base_means, b.base_stds, b.base_signal_lens = s.kmer[:n], s.means[:n], s.stds[:n], s.lens[:n]
        b.signals, b.labels = s.signals[:n], s.labels[:n]
        b.info_text, b.info_off = s.info_text, s.info_off
        return b

    def _iter_mapped(self, L):
        """Plain files are parsed in place from a read-only mapping: no read() copy, and the page-cache
        faults are taken by the parser's worker threads instead of one reading thread."""
        import mmap
        with open(self.path, "rb") as f:
            size = os.fstat(f.fileno()).st_size
            start, end = shard_bounds(f, *self.byte_range) if self.byte_range is not None else (0, size)
            if end <= start:
                return
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            view = np.frombuffer(mm, np.uint8)
            base = view.ctypes.data
            est_line = 4 * self.T * 10 + self.T * self.S * 10       # refined after the first block
            pos, k = start, 0
            while pos < end:
                want = min(int(est_line * self.batch_sites * 1.05) + (1 << 16), end - pos)
                final = pos + want == end
                if final and bytes(view[pos:min(pos + 4096, end)]).isspace() and bytes(view[pos:end]).isspace():
                    return
                s = self._slot(k % self.nslots)
                n, used = self._parse(L, s, base + pos, want, final)
                if n == 0:
                    if final:
                        return
                    est_line *= 2                        # a line longer than the whole block: look further
                    continue
                pos += used
                est_line = max(64, used // n)
                self.sites_read += n
                k += 1
                yield self._batch(s, n, k - 1)
        finally:
            del view
            try:
                mm.close()
            except BufferError:                          # a caller still holds a view; the mapping goes with it
                pass

    def __iter__(self):
        L = _native.lib()
        if not self.path.endswith(".gz"):
            yield from self._iter_mapped(L)
            return
        f = gzip.open(self.path, "rb") if self.path.endswith(".gz") else open(self.path, "rb")
        remaining = None
        with f:
            if self.byte_range is not None:
                start, end = shard_bounds(f, *self.byte_range)
                f.seek(start)
                remaining = end - start
            if remaining is None and not self.path.endswith(".gz"):
                remaining = os.path.getsize(self.path)              # known for plain files: never over-allocate
            eof = False
            est_line = 4 * self.T * 10 + self.T * self.S * 10       # refined after the first block
            k = 0
            arr = bytearray(0)                                      # read buffer, reused; arr[:fill] is valid
            fill = 0
            while True:
                want = int(est_line * self.batch_sites * 1.05) + (1 << 16)
                if remaining is not None:
                    want = min(want, fill + remaining + 1)
                if len(arr) < want:                       # grow (calloc-backed: untouched pages cost nothing)
                    bigger = bytearray(want)
                    bigger[:fill] = arr[:fill]
                    arr = bigger
                mv = memoryview(arr)
                while not eof and fill < want:
                    ask = want - fill if remaining is None else min(want - fill, remaining)
                    got = f.readinto(mv[fill:fill + ask]) if ask > 0 else 0
                    if not got:
                        eof = True
                        break
                    if remaining is not None:
                        remaining -= got
                    fill += got
                if fill == 0 or bytes(mv[:min(fill, 4096)]).isspace() and bytes(mv[:fill]).isspace():
                    return
                s = self._slot(k % self.nslots)
                n, used = C.c_int64(0), C.c_int64(0)
                base = (C.c_char * len(arr)).from_buffer(arr)
                while True:
                    rc = L.dsp_parse_features(
                        base, fill, int(eof), self.T, self.S, self.batch_sites,
                        s.kmer.data_ptr(), s.means.data_ptr(), s.stds.data_ptr(), s.lens.data_ptr(), s.signals.data_ptr(),
                        s.labels.data_ptr(), s.info_text.ctypes.data, s.info_text.size, s.info_off.ctypes.data,
                        C.byref(n), C.byref(used), self.nthreads)
                    if rc == 4 and s.info_text.size < fill:       # DSP_ERR_NOMEM: unusually long sample-info columns
                        s.info_text = np.empty(min(fill, s.info_text.size * 4), np.uint8)
                        continue
                    _native.check(rc, "dsp_parse_features(%s)" % self.path)
                    break
                del base
                n, used = int(n.value), int(used.value)
                if n == 0:
                    if eof:
                        return
                    est_line *= 2                        # a line longer than the whole block: read more
                    mv.release()
                    continue
                b = FeatureBatch()
                b.n, b.seq_len, b.slot = n, self.T, k % self.nslots
                b.kmer, b.base_means, b.base_stds, b.base_signal_lens = s.kmer[:n], s.means[:n], s.stds[:n], s.lens[:n]
                b.signals, b.labels = s.signals[:n], s.labels[:n]
                b.info_text, b.info_off = s.info_text, s.info_off
                mv[:fill - used] = mv[used:fill]         # carry the (small) unparsed tail to the front
                fill -= used
                mv.release()
                est_line = max(64, used // n)
                self.sites_read += n
                k += 1
                yield b


def format_calls(batch, probs, labels, nthreads=None, as_array=False):
    """bytes of the call_mods lines of one batch (``call_modifications.py:175-188``), each
    terminated by a newline (what ``_write_predstr_to_file`` writes, ``:279-280``).
    ``probs`` (n, 2) float32 and ``labels`` (n) int32: numpy arrays or CPU torch tensors.
    ``as_array``: return the uint8 array the formatter wrote into instead of a bytes copy of it."""
    L = _native.lib()
    n = batch.n
    p = np.ascontiguousarray(np.asarray(probs, dtype=np.float32))
    lab = np.ascontiguousarray(np.asarray(labels, dtype=np.int32))
    if p.shape != (n, 2) or lab.shape != (n,):
        raise ValueError("format_calls: probs must be (n, 2) and labels (n,) for the batch's n = %d" % n)
    kmer = batch.kmer if isinstance(batch.kmer, np.ndarray) else batch.kmer.numpy()
    kmer = np.ascontiguousarray(kmer[:n], dtype=np.float32)
    cap = int(batch.info_off[n]) + n * (64 + batch.seq_len)
    out = np.empty(max(cap, 1), np.uint8)                  # not zero-filled; the formatter's threads touch the pages
    used = C.c_int64(0)
    _native.check(L.dsp_format_calls(batch.info_text.ctypes.data, batch.info_off.ctypes.data, kmer.ctypes.data,
                                     batch.seq_len, p.ctypes.data, lab.ctypes.data, n,
                                     out.ctypes.data, cap, C.byref(used), int(nthreads or min(16, os.cpu_count() or 1))),
                  "dsp_format_calls")
    return out[:used.value] if as_array else out[:used.value].tobytes()
