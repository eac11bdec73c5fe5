"""Binary feature hand-off between ``extract`` and ``call_mods``.

The reference couples its two commands through a 12-column text file (``_features_to_str``,
``deepsignal_plant/extract_features.py:381-395``; read back line by line in
``call_modifications.py:55-127``): 2.1 KB of decimal text for the 1 040 bytes of float32 that
``ModelBiLSTM.forward`` consumes per site, and the parser -- not the GPU -- bounds ``call_mods``.
A ``.dspf`` file holds the same information in the form the model takes it:

    file   = header (64 B) , block , block , ...
    header = b"DSPFEAT1" , u32 byte-order mark 0x01020304 , u32 seq_len , u32 signal_len , zero padding
    block  = block header (64 B) , sections (each padded to 64 B)
    block header = b"DSPFBLK1" , u64 n sites , u64 info_bytes , u64 bytes of the whole block , zero padding
    sections     = kmer f32[n,T] , base_means f32[n,T] , base_stds f32[n,T] , base_signal_lens f32[n,T] ,
                   signals f32[n,T,S] , labels i32[n] , info_off i64[n+1] , info_text u8[info_bytes]

``info_text[info_off[i]:info_off[i+1]]`` = the six leading columns of site i's line joined by tabs (what the
call_mods output line starts with, ``call_modifications.py:89,175-188``).  Values are float32(the float64 the
text file would print), i.e. exactly what ``FloatTensor(list of float(text))`` gives the model, so the calls
written from a ``.dspf`` file equal the calls written from the text file byte for byte (tests/test_feature_bin.py).

``FeatureBinReader`` yields the same ``feature_io.FeatureBatch`` objects as the text reader: every section of a
batch is read straight into its page-locked slot by ``os.preadv`` on a few host threads (one copy, page cache ->
pinned memory; the calls release the GIL).  Ranks of a torchrun job take contiguous site ranges.
"""
from __future__ import annotations

import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .feature_io import FeatureBatch, _Slot

FILE_MAGIC = b"DSPFEAT1"
BLOCK_MAGIC = b"DSPFBLK1"
BOM = 0x01020304
HEADER_BYTES = 64
ALIGN = 64
SUFFIX = ".dspf"
PIECE = 8 << 20                    # bytes per preadv call: pieces of one section go to different threads


def _pad(nbytes):
    return (nbytes + ALIGN - 1) & ~(ALIGN - 1)


def block_layout(n, T, S, info_bytes):
    """Byte offsets of a block's sections relative to the block start -> (dict, block bytes)."""
    off = HEADER_BYTES
    secs = {}
    for name, nbytes in (("kmer", 4 * n * T), ("means", 4 * n * T), ("stds", 4 * n * T), ("lens", 4 * n * T),
                         ("signals", 4 * n * T * S), ("labels", 4 * n), ("info_off", 8 * (n + 1)), ("info_text", info_bytes)):
        secs[name] = off
        off += _pad(nbytes)
    return secs, off


def is_feature_bin(path):
    """True when ``path`` starts with the ``.dspf`` magic."""
    try:
        with open(path, "rb") as f:
            return f.read(8) == FILE_MAGIC
    except OSError:
        return False


class FeatureBinWriter:
    """Write a ``.dspf`` file block by block (a block = whatever one ``write`` call brings)."""

    def __init__(self, path, seq_len, signal_len):
        self.T, self.S = int(seq_len), int(signal_len)
        self.sites = 0
        self._f = open(path, "wb")
        self._f.write(FILE_MAGIC + struct.pack("<III", BOM, self.T, self.S) + bytes(HEADER_BYTES - 20))

    def write(self, kmer, base_means, base_stds, base_signal_lens, signals, labels, info_text, info_off):
        """One block.  Arrays as ``ModelBiLSTM.forward`` takes them (anything ``np.asarray`` reads; float32), ``labels``
        (n,) or a scalar, ``info_text`` uint8 / bytes with ``info_off`` (n + 1) offsets into it."""
        T, S = self.T, self.S
        f32 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float32)
        kmer, means, stds, lens, signals = f32(kmer), f32(base_means), f32(base_stds), f32(base_signal_lens), f32(signals)
        n = kmer.shape[0]
        if n == 0:
            return
        for a in (kmer, means, stds, lens):
            if a.shape != (n, T):
                raise ValueError("feature arrays must be (n, %d), got %s" % (T, a.shape))
        if signals.shape != (n, T, S):
            raise ValueError("signals must be (n, %d, %d), got %s" % (T, S, signals.shape))
        labels = np.ascontiguousarray(np.broadcast_to(np.asarray(labels, dtype=np.int32), (n,)))
        off = np.ascontiguousarray(np.asarray(info_off, dtype=np.int64)[:n + 1])
        if off.shape != (n + 1,) or np.any(np.diff(off) < 0):
            raise ValueError("info_off must hold n + 1 non-decreasing offsets")
        text = np.frombuffer(info_text, np.uint8) if isinstance(info_text, (bytes, bytearray, memoryview)) else np.asarray(info_text, np.uint8)
        text = np.ascontiguousarray(text[int(off[0]):int(off[n])])
        off = off - off[0]
        secs, total = block_layout(n, T, S, text.size)
        w = self._f.write
        w(BLOCK_MAGIC + struct.pack("<QQQ", n, text.size, total) + bytes(HEADER_BYTES - 32))
        for a in (kmer, means, stds, lens, signals, labels, off, text):
            w(memoryview(a).cast("B") if a.size else b"")
            w(bytes(_pad(a.nbytes) - a.nbytes))
        self.sites += n

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_header(f):
    """(seq_len, signal_len) from an open binary file; raises ValueError on anything that is not a ``.dspf`` file."""
    f.seek(0)
    h = f.read(HEADER_BYTES)
    if len(h) < HEADER_BYTES or h[:8] != FILE_MAGIC:
        raise ValueError("not a binary feature file (bad magic)")
    bom, T, S = struct.unpack("<III", h[8:20])
    if bom != BOM:
        raise ValueError("binary feature file written with another byte order")
    return int(T), int(S)


def scan_blocks(path):
    """-> (seq_len, signal_len, [(file offset, n sites, info_bytes), ...]) by hopping from block header to block header."""
    blocks = []
    with open(path, "rb") as f:
        T, S = read_header(f)
        size = f.seek(0, os.SEEK_END)
        pos = HEADER_BYTES
        while pos < size:
            f.seek(pos)
            h = f.read(32)
            if len(h) < 32 or h[:8] != BLOCK_MAGIC:
                raise ValueError("%s: damaged block header at byte %d" % (path, pos))
            n, info_bytes, total = struct.unpack("<QQQ", h[8:32])
            if total != block_layout(n, T, S, info_bytes)[1] or pos + total > size:
                raise ValueError("%s: truncated or inconsistent block at byte %d" % (path, pos))
            blocks.append((pos, int(n), int(info_bytes)))
            pos += total
    return T, S, blocks


class FeatureBinReader:
    """Iterate ``FeatureBatch`` objects over a ``.dspf`` file; same slot recycling rule as ``FeatureFileReader``.
    ``site_range=(lo, hi)`` restricts the reader to sites [lo, hi) of the file (contiguous shards, one per rank)."""

    def __init__(self, path, seq_len=None, signal_len=None, batch_sites=65536, pinned=None, slots=4, nthreads=None,
                 site_range=None):
        import torch
        self.path = path
        self.T, self.S, self.blocks = scan_blocks(path)
        if seq_len is not None and int(seq_len) != self.T or signal_len is not None and int(signal_len) != self.S:
            raise ValueError("%s holds %d-mers with %d samples per base, not --seq_len %s --signal_len %s"
                             % (path, self.T, self.S, seq_len, signal_len))
        self.batch_sites = int(batch_sites)
        self.pinned = torch.cuda.is_available() if pinned is None else bool(pinned)
        self.nthreads = max(1, min(int(nthreads or 8), 16))
        self.nslots = int(slots)
        self._slots = [None] * self.nslots
        self.total_sites = sum(b[1] for b in self.blocks)
        lo, hi = (0, self.total_sites) if site_range is None else site_range
        self.site_range = (max(0, int(lo)), min(self.total_sites, int(hi)))
        self.sites_read = 0

    def _slot(self, i):
        if self._slots[i] is None:
            self._slots[i] = _Slot(self.batch_sites, self.T, self.S, self.pinned)
        return self._slots[i]

    def _spans(self):
        """(block offset, block n, info_bytes, a, z): site ranges [a, z) inside blocks, at most ``batch_sites`` long,
        covering ``site_range`` in file order."""
        lo, hi = self.site_range
        g = 0
        for off, n, info_bytes in self.blocks:
            a0, z0 = max(lo - g, 0), min(hi - g, n)
            for a in range(a0, z0, self.batch_sites):
                yield off, n, info_bytes, a, min(a + self.batch_sites, z0)
            g += n
            if g >= hi:
                return

    def __iter__(self):
        fd = os.open(self.path, os.O_RDONLY)
        pool = ThreadPoolExecutor(self.nthreads)
        T, S = self.T, self.S

        def fill(dst, offset):
            # dst: writable byte view; preadv may return short counts on signals / huge requests
            done, want = 0, len(dst)
            while done < want:
                got = os.preadv(fd, [dst[done:]], offset + done)
                if got <= 0:
                    raise IOError("%s: unexpected end of file at byte %d" % (self.path, offset + done))
                done += got

        try:
            k = 0
            for off, n, info_bytes, a, z in self._spans():
                m = z - a
                s = self._slot(k % self.nslots)
                secs, _ = block_layout(n, T, S, info_bytes)
                io = s.info_off[:m + 1]
                fill(memoryview(io).cast("B"), off + secs["info_off"] + 8 * a)
                t0, t1 = int(io[0]), int(io[m])
                if t1 < t0 or t1 > info_bytes:
                    raise ValueError("%s: damaged sample-info offsets in the block at byte %d" % (self.path, off))
                if s.info_text.size < t1 - t0:
                    s.info_text = np.empty(int((t1 - t0) * 1.25) + 64, np.uint8)
                jobs = []

                def add(arr_bytes, offset):
                    for p in range(0, len(arr_bytes), PIECE):
                        jobs.append(pool.submit(fill, arr_bytes[p:p + PIECE], offset + p))

                row = 4 * T
                for name, t in (("kmer", s.kmer), ("means", s.means), ("stds", s.stds), ("lens", s.lens)):
                    add(memoryview(t.numpy()).cast("B")[:m * row], off + secs[name] + a * row)
                add(memoryview(s.signals.numpy()).cast("B")[:m * row * S], off + secs["signals"] + a * row * S)
                add(memoryview(s.labels.numpy()).cast("B")[:4 * m], off + secs["labels"] + 4 * a)
                if t1 > t0:
                    add(memoryview(s.info_text)[:t1 - t0], off + secs["info_text"] + t0)
                io -= t0
                for j in jobs:
                    j.result()
                b = FeatureBatch()
                b.n, b.seq_len, b.slot = m, T, k % self.nslots
                b.kmer, b.base_means, b.base_stds, b.base_signal_lens = s.kmer[:m], s.means[:m], s.stds[:m], s.lens[:m]
                b.signals, b.labels = s.signals[:m], s.labels[:m]
                b.info_text, b.info_off = s.info_text, s.info_off
                self.sites_read += m
                k += 1
                yield b
        finally:
            pool.shutdown(wait=True)
            os.close(fd)


def pack_feature_file(text_path, bin_path, seq_len=None, signal_len=None, batch_sites=65536, nthreads=None):
    """The reference's text feature file -> ``.dspf`` (``dsp_parse_features`` does the parsing).  Returns the site count."""
    from .feature_io import FeatureFileReader
    reader = FeatureFileReader(text_path, seq_len, signal_len, batch_sites=batch_sites, pinned=False, slots=2, nthreads=nthreads)
    with FeatureBinWriter(bin_path, reader.T, reader.S) as w:
        for b in reader:
            w.write(b.kmer.numpy(), b.base_means.numpy(), b.base_stds.numpy(), b.base_signal_lens.numpy(), b.signals.numpy(),
                    b.labels.numpy(), b.info_text, b.info_off[:b.n + 1])
        return w.sites
