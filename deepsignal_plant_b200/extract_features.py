"""Feature extraction from decoded re-squiggled reads on the GPU (``dsp_extract_features``).

Mirrors the numeric part of ``deepsignal_plant/extract_features.py`` of the reference:
``_rescale_signals`` (``:276-277``), ``_normalize_signals`` (``:179-190``, method ``mad``), the per-site
body of ``_extract_features`` (``:318-372``) and ``_get_signals_rect`` (``:232-251``), and hands the
result over either in the reference's own shape (``_extract_features`` below returns the same
12-tuples, ``:370-372``) or as the five device tensors ``ModelBiLSTM.forward`` takes, without the
2.1 KB-per-site text detour of the feature file.

What is NOT here: reading fast5 files.  h5py is absent in this image, so the entry points start at
what ``_get_alignment_info_from_fast5`` / ``_get_label_raw`` / ``_get_scaling_of_a_read``
(``:150-176,37-91,255-273``) return -- one dict per read (see ``pack_reads``).  A maintainer with h5py
wraps those three accessors around ``pack_reads``; INTEGRATION.md shows the stub.

The arithmetic runs in libdsp_b200 only; there is no CPU fallback (``pack_reads`` / ``find_sites`` are
host-side index bookkeeping, not arithmetic).  Both ``normalize_method`` values of the reference run on
the device (``mad``, the default, and ``zscore``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native

base2code_dna = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'N': 4, 'W': 5, 'S': 6, 'M': 7, 'K': 8, 'R': 9,
                 'Y': 10, 'B': 11, 'V': 12, 'D': 13, 'H': 14, 'Z': 15}
iupac_alphabets = {'A': ['A'], 'T': ['T'], 'C': ['C'], 'G': ['G'], 'R': ['A', 'G'], 'M': ['A', 'C'],
                   'S': ['C', 'G'], 'Y': ['C', 'T'], 'K': ['G', 'T'], 'W': ['A', 'T'], 'B': ['C', 'G', 'T'],
                   'D': ['A', 'G', 'T'], 'H': ['A', 'C', 'T'], 'V': ['A', 'C', 'G'], 'N': ['A', 'C', 'G', 'T']}
key_sep = "||"
_ALPHABET = np.frombuffer(b"ACGTNWSMKRYBVDHZ", np.uint8)
_ALPHABET_LUT = np.zeros(256, bool)
_ALPHABET_LUT[_ALPHABET] = True


def get_motif_seqs(motifs, is_dna=True):
    """``utils/process_utils.py:115-147``: expand comma-separated IUPAC motifs (DNA only here)."""
    if not is_dna:
        raise ValueError("only DNA motifs are supported")
    out = []
    for ori in motifs.strip().split(','):
        seqs = ['']
        for b in ori.strip().upper():
            seqs = [s + x for s in seqs for x in iupac_alphabets[b]]
        out += seqs
    return out


_STAGING = {}          # dtype -> [pinned tensor, event of the last copy out of it]


def _upload(a, device):
    """Host array -> device tensor.  Large arrays (views of a memory-mapped archive, typically) go through a
    reusable page-locked staging buffer: a plain memcpy out of the page cache, then an asynchronous copy at
    full PCIe rate -- a pageable cudaMemcpy straight from a file mapping runs at about 1 GB/s."""
    a = np.ascontiguousarray(a)
    if a.nbytes < (1 << 20):
        return torch.from_numpy(np.array(a)).to(device)
    tdtype = torch.from_numpy(np.zeros(0, a.dtype)).dtype
    slot = _STAGING.get(a.dtype)
    if slot is None or slot[0].numel() < a.size:
        slot = _STAGING[a.dtype] = [torch.empty(int(a.size * 1.25), dtype=tdtype).pin_memory(), None]
    if slot[1] is not None:
        slot[1].synchronize()                                   # the previous copy out of this buffer has finished
    np.copyto(slot[0][:a.size].numpy(), a.reshape(-1))
    with torch.cuda.device(device):
        out = slot[0][:a.size].to(device, non_blocking=True).reshape(a.shape)
        slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream())
    return out


class ReadBatch:
    """Flat arrays of a batch of decoded reads (host numpy; ``to_device`` uploads them once)."""

    def __init__(self, reads):
        n = len(reads)
        self.n_reads = n
        self.readname = [r["readname"] for r in reads]
        self.strand = [r["strand"] for r in reads]
        self.alignstrand = [r["alignstrand"] for r in reads]
        self.chrom = [r["chrom"] for r in reads]
        self.chrom_start = np.array([r["chrom_start"] for r in reads], np.int64)
        self.raw_off = np.zeros(n + 1, np.int64)
        self.ev_off = np.zeros(n + 1, np.int64)
        for i, r in enumerate(reads):
            self.raw_off[i + 1] = self.raw_off[i] + len(r["raw"])
            self.ev_off[i + 1] = self.ev_off[i] + len(r["ev_len"])
        cat = lambda key, dt: (np.concatenate([np.asarray(r[key]) for r in reads]).astype(dt, copy=False)
                               if n else np.zeros(0, dt))
        self.raw = cat("raw", np.int16)
        self.ev_start = cat("ev_start", np.int64)
        self.ev_len = cat("ev_len", np.int64)
        self.ev_base = np.frombuffer("".join(r["ev_base"] for r in reads).encode("ascii"), np.uint8).copy()
        # _get_scaling_of_a_read returns (None, None) when the channel info cannot be read (:271-273)
        self.scaling = np.array([np.nan if r.get("scaling") is None else r["scaling"] for r in reads], np.float64)
        self.offset = np.array([0.0 if r.get("scaling") is None else r["offset"] for r in reads], np.float64)
        self._dev = None
        self._validate()

    def bad_reads(self):
        """Boolean mask of the reads the reference's per-read ``try/except`` would count as errors and skip
        (``_extract_features``: ``error += 1``, extract_features.py:373-375): a base outside the alphabet
        (``base2code_dna[x]`` raises KeyError) or an event that reaches outside the read's raw signal."""
        n = self.n_reads
        bad = np.zeros(n, bool)
        if n == 0 or self.ev_len.shape[0] == 0:
            return bad
        read_of_ev = np.repeat(np.arange(n), np.diff(self.ev_off))
        ok = _ALPHABET_LUT[self.ev_base]
        if not ok.all():
            bad[np.unique(read_of_ev[~ok])] = True
        neg = (self.ev_start < 0) | (self.ev_len < 0)
        if neg.any():
            bad[np.unique(read_of_ev[neg])] = True
        over = (self.ev_start + self.ev_len) > np.diff(self.raw_off)[read_of_ev]
        if over.any():
            bad[np.unique(read_of_ev[over])] = True
        return bad

    def drop_bad_reads(self):
        """-> (batch without the offending reads, number dropped); one bad read no longer aborts the whole batch."""
        bad = self.bad_reads()
        nbad = int(bad.sum())
        if nbad == 0:
            return self, 0
        keep = np.flatnonzero(~bad)
        parts = [self.slice(int(i), int(i) + 1) for i in keep]
        b = object.__new__(ReadBatch)
        b.n_reads = len(keep)
        for k in ("readname", "strand", "alignstrand", "chrom"):
            src = getattr(self, k)
            setattr(b, k, [src[int(i)] for i in keep] if isinstance(src, list) else src[keep])
        b.chrom_start, b.scaling, b.offset = self.chrom_start[keep], self.scaling[keep], self.offset[keep]
        b.raw_off = np.concatenate([[0], np.cumsum([p.raw.shape[0] for p in parts])]).astype(np.int64)
        b.ev_off = np.concatenate([[0], np.cumsum([p.ev_len.shape[0] for p in parts])]).astype(np.int64)
        cat = lambda f, dt: np.concatenate([getattr(p, f) for p in parts]).astype(dt, copy=False) if parts else np.zeros(0, dt)
        b.raw, b.ev_start, b.ev_len, b.ev_base = cat("raw", np.int16), cat("ev_start", np.int64), cat("ev_len", np.int64), cat("ev_base", np.uint8)
        b._dev = None
        return b, nbad

    def _validate(self):
        if self.ev_base.shape[0] != self.ev_len.shape[0] or self.ev_start.shape[0] != self.ev_len.shape[0]:
            raise ValueError("event columns differ in length")         # the reference asserts (:87-88)
        bad = self.bad_reads()
        if bad.any():
            i = int(np.argmax(bad))
            lo, hi = int(self.ev_off[i]), int(self.ev_off[i + 1])
            ok = _ALPHABET_LUT[self.ev_base[lo:hi]]
            if not ok.all():
                raise KeyError(chr(int(self.ev_base[lo + int(np.argmin(ok))])))  # base2code_dna[x] in the reference
            raise ValueError("an event reaches outside its read's raw signal")

    # ---- the archive form: flat arrays in an .npz (what a fast5 decoder writes once; see save_reads)
    ARCHIVE_KEYS = ("raw", "raw_off", "ev_off", "scaling", "offset", "ev_start", "ev_len", "ev_base",
                    "readname", "strand", "alignstrand", "chrom", "chrom_start")

    def arrays(self):
        return dict(raw=self.raw, raw_off=self.raw_off, ev_off=self.ev_off, scaling=self.scaling, offset=self.offset,
                    ev_start=self.ev_start, ev_len=self.ev_len, ev_base=self.ev_base,
                    readname=np.array(self.readname), strand=np.array(self.strand),
                    alignstrand=np.array(self.alignstrand), chrom=np.array(self.chrom), chrom_start=self.chrom_start)

    def slice(self, lo, hi):
        """Reads [lo, hi) as a batch of their own (views of the flat arrays, offsets rebased)."""
        b = object.__new__(ReadBatch)
        a, z = int(self.raw_off[lo]), int(self.raw_off[hi])
        c, d = int(self.ev_off[lo]), int(self.ev_off[hi])
        b.n_reads = hi - lo
        for k in ("readname", "strand", "alignstrand", "chrom"):
            setattr(b, k, getattr(self, k)[lo:hi])
        b.chrom_start, b.scaling, b.offset = self.chrom_start[lo:hi], self.scaling[lo:hi], self.offset[lo:hi]
        b.raw_off, b.ev_off = self.raw_off[lo:hi + 1] - a, self.ev_off[lo:hi + 1] - c
        b.raw, b.ev_start, b.ev_len, b.ev_base = self.raw[a:z], self.ev_start[c:d], self.ev_len[c:d], self.ev_base[c:d]
        b._dev = None
        return b

    def to_device(self, device):
        if self._dev is None or self._dev["device"] != device:
            up = lambda a: _upload(a, device)
            self._dev = dict(device=device, raw=up(self.raw), raw_off=up(self.raw_off), scaling=up(self.scaling),
                             offset=up(self.offset), ev_start=up(self.ev_start), ev_len=up(self.ev_len),
                             ev_base=up(self.ev_base), ev_off=up(self.ev_off))
        return self._dev


def save_reads(path, reads):
    """Write decoded reads as an .npz archive of flat arrays (``ReadBatch.ARCHIVE_KEYS``): the input
    ``call_mods --input_path reads.npz`` takes in place of a fast5 directory."""
    np.savez(path, **pack_reads(reads).arrays())


class _MappedArchive:
    """The members of an uncompressed .npz (``np.savez`` stores them as plain .npy files inside the zip)
    memory-mapped in place: loading a 600 MB archive costs page-table entries instead of a 600 MB copy,
    and the chunks go from the page cache straight into the host->device copy."""

    def __init__(self, path):
        import zipfile
        self.files, self._arr = [], {}
        with zipfile.ZipFile(path) as zf, open(path, "rb") as f:
            for info in zf.infolist():
                name = info.filename[:-4] if info.filename.endswith(".npy") else info.filename
                self.files.append(name)
                if info.compress_type != zipfile.ZIP_STORED:
                    self._arr[name] = None                           # compressed member: read it the ordinary way
                    continue
                f.seek(info.header_offset)
                hdr = f.read(30)                                     # local file header: name and extra lengths at 26, 28
                start = info.header_offset + 30 + int.from_bytes(hdr[26:28], "little") + int.from_bytes(hdr[28:30], "little")
                f.seek(start)
                version = np.lib.format.read_magic(f)
                shape, fortran, dtype = (np.lib.format.read_array_header_1_0(f) if version == (1, 0)
                                         else np.lib.format.read_array_header_2_0(f))
                if dtype.hasobject or fortran:
                    self._arr[name] = None
                    continue
                n = int(np.prod(shape))
                self._arr[name] = (np.memmap(path, dtype=dtype, mode="c", offset=f.tell(), shape=shape) if n
                                   else np.zeros(shape, dtype))
        self._path = path

    def __getitem__(self, k):
        a = self._arr[k]
        if a is None:
            a = self._arr[k] = np.load(self._path)[k]
        return a


def load_reads(path):
    """-> ReadBatch straight from the archive's flat arrays (memory-mapped, no per-read objects); reads the
    reference would count as errors are dropped (``n_errors``)."""
    z = _MappedArchive(path)
    missing = [k for k in ReadBatch.ARCHIVE_KEYS if k not in z.files]
    if missing:
        raise ValueError("%s is not a decoded-reads archive (missing %s)" % (path, ", ".join(missing)))
    b = object.__new__(ReadBatch)
    b.n_reads = int(z["readname"].shape[0])
    for k in ("readname", "strand", "alignstrand", "chrom"):
        setattr(b, k, np.asarray(z[k]).tolist())
    b.chrom_start = np.asarray(z["chrom_start"]).astype(np.int64, copy=False)
    b.raw_off, b.ev_off = z["raw_off"].astype(np.int64, copy=False), z["ev_off"].astype(np.int64, copy=False)
    b.raw = z["raw"].astype(np.int16, copy=False)
    b.ev_start, b.ev_len = z["ev_start"].astype(np.int64, copy=False), z["ev_len"].astype(np.int64, copy=False)
    b.ev_base = z["ev_base"].astype(np.uint8, copy=False)
    b.scaling, b.offset = z["scaling"].astype(np.float64, copy=False), z["offset"].astype(np.float64, copy=False)
    b._dev = None
    if b.ev_base.shape[0] != b.ev_len.shape[0] or b.ev_start.shape[0] != b.ev_len.shape[0]:
        raise ValueError("event columns differ in length")
    # the reference wraps every read in try/except and only counts the failures (extract_features.py:373-375):
    # a read with a base outside the alphabet or an event outside its raw signal is skipped, not fatal
    b, nbad = b.drop_bad_reads()
    if nbad:
        print("extract_features: %d of %d reads skipped (error)" % (nbad, b.n_reads + nbad))
    b.n_errors = nbad
    return b


def _read_position_file(position_file):
    """``extract_features.py:519-528``: set of ``chrom||pos||strand`` keys."""
    positions = set()
    with open(position_file, "r") as rf:
        for line in rf:
            words = line.strip().split("\t")
            if len(words) < 3:
                raise ValueError("--position file in wrong format. "
                                 "If you didn't use Tab as delimiter, Please do.")
            positions.add(key_sep.join(words[:3]))
    return positions


def parse_region_str(regionstr):
    """``utils/process_utils.py:163-187``: ``chrom:start-end`` | ``chrom:start`` | ``chrom`` (0-based, half-open)."""
    try:
        if regionstr is None:
            return None, None, None
        if ":" in regionstr:
            chrom, se = regionstr.strip().split(":")
            if "-" in se:
                s_, e_ = se.split("-")
                return chrom, int(s_), int(e_)
            return chrom, int(se), None
        return regionstr.strip(), None, None
    except Exception:
        raise ValueError("--region not set right!")


def get_contig2len(ref_path):
    """``utils/ref_reader.py:7-30``: contig name (up to the first blank) -> sequence length of a FASTA."""
    chrom2len, name, n = {}, "", 0
    with open(ref_path, "r") as rf:
        for line in rf:
            if line.startswith(">"):
                if name != "" and n:
                    chrom2len[name] = n
                name, n = line.strip()[1:].split(" ")[0], 0
            else:
                n += len(line.strip())
    chrom2len[name] = n
    return chrom2len


def pack_reads(reads):
    """reads: list of dicts with readname, strand ('t'/'c'), alignstrand, chrom, chrom_start
    (``_get_alignment_info_from_fast5``), raw (int16 DAC samples), ev_start (already shifted by
    ``read_start_rel_to_raw``), ev_len, ev_base (``_get_label_raw``), scaling, offset
    (``_get_scaling_of_a_read``; None = no channel info)."""
    return reads if isinstance(reads, ReadBatch) else ReadBatch(reads)


class Sites:
    """Which (read, base) pairs to extract, and their coordinates (``:341-355``)."""
    __slots__ = ("site_read", "site_ev", "pos", "pos_in_strand", "_dev")

    def __init__(self, site_read, site_ev, pos, pos_in_strand):
        self.site_read, self.site_ev, self.pos, self.pos_in_strand = site_read, site_ev, pos, pos_in_strand
        self._dev = None

    def __len__(self):
        return int(self.site_read.shape[0])

    def to_device(self, device):
        if self._dev is None or self._dev[0] != device:
            self._dev = (device, torch.from_numpy(self.site_read).to(device), torch.from_numpy(self.site_ev).to(device))
        return self._dev[1], self._dev[2]


def find_sites(batch, motif_seqs, methyloc, chrom2len, kmer_len, positions=None, regioninfo=(None, None, None)):
    """``get_refloc_of_methysite_in_motif`` (``utils/process_utils.py:97-112``) over every read of the
    batch at once, then the reference's site filters in its order (``:341-355``): margin of
    ``(kmer_len-1)//2`` bases, strand-aware position, region, position set.  Index bookkeeping only."""
    if kmer_len % 2 == 0:
        raise ValueError("kmer_len must be odd")
    num_bases = (kmer_len - 1) // 2
    motifs = sorted(set(motif_seqs))
    mlen = len(motifs[0])
    rg_chrom, rg_start, rg_end = regioninfo
    nb = batch.ev_base.shape[0]
    empty = Sites(np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64))
    if nb < mlen:
        return empty
    hit = np.zeros(nb - mlen + 1, bool)
    for m in motifs:
        mb = np.frombuffer(m.encode("ascii"), np.uint8)
        h = batch.ev_base[0:nb - mlen + 1] == mb[0]
        for k in range(1, mlen):
            h &= batch.ev_base[k:nb - mlen + 1 + k] == mb[k]
        hit |= h
    start = np.nonzero(hit)[0]
    rd = np.searchsorted(batch.ev_off, start, side="right") - 1
    rlen = np.diff(batch.ev_off)[rd]
    loc0 = start - batch.ev_off[rd]
    ok = loc0 + mlen <= rlen                                  # the motif must not run into the next read
    loc = loc0 + methyloc
    ok &= (loc >= num_bases) & (loc < rlen - num_bases)
    rd, loc, rlen = rd[ok], loc[ok], rlen[ok]
    cstart = batch.chrom_start[rd]
    minus = np.array([s == '-' for s in batch.alignstrand], bool)[rd] if batch.n_reads else np.zeros(0, bool)
    pos = np.where(minus, cstart + rlen - 1 - loc, cstart + loc)
    if chrom2len is not None:
        clen = np.array([chrom2len.get(c, -1) for c in batch.chrom], np.int64)[rd]
        pis = np.where(clen < 0, -1, np.where(minus, clen - 1 - pos, pos))
    else:
        pis = np.full(pos.shape, -1, np.int64)
    keep = np.ones(pos.shape, bool)
    if rg_chrom is not None:
        same = np.array([c == rg_chrom for c in batch.chrom], bool)[rd]
        rs = cstart if rg_start is None else np.full(pos.shape, rg_start, np.int64)
        re_ = cstart + rlen if rg_end is None else np.full(pos.shape, rg_end, np.int64)
        keep &= same & ~((rs >= cstart + rlen) | (re_ <= cstart)) & (pos >= rs) & (pos < re_)
    if positions is not None:
        keep &= np.array([key_sep.join([batch.chrom[r], str(int(p)), batch.alignstrand[r]]) in positions
                          for r, p in zip(rd, pos)], bool)
    rd, loc, pos, pis = rd[keep], loc[keep], pos[keep], pis[keep]
    return Sites(rd.astype(np.int32), (batch.ev_off[rd] + loc).astype(np.int64), pos.astype(np.int64), pis.astype(np.int64))


def find_sites_device(batch, motif_seqs, methyloc, chrom2len, kmer_len, positions=None, regioninfo=(None, None, None),
                      device=None):
    """``find_sites`` evaluated by ``dsp_find_sites`` on the batch's device-resident event table: same sites,
    same order; the result also keeps its device copies, so ``extract_tensors`` uploads nothing.  A
    ``positions`` set (a host-side string-set lookup, ``:354``) falls back to ``find_sites``."""
    if positions is not None:
        return find_sites(batch, motif_seqs, methyloc, chrom2len, kmer_len, positions, regioninfo)
    if kmer_len % 2 == 0:
        raise ValueError("kmer_len must be odd")
    if not torch.cuda.is_available():
        raise _native.DspError("dsp_find_sites needs a CUDA device; there is no CPU fallback")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    motifs = sorted(set(motif_seqs))
    mlen = len(motifs[0])
    if any(len(m) != mlen for m in motifs):
        raise ValueError("motifs must have one length")             # the reference takes len(list(motifset)[0]), :107
    L = _native.lib()
    d = batch.to_device(device)
    n_reads, n_events = batch.n_reads, int(batch.ev_base.shape[0])
    rg_chrom, rg_start, rg_end = regioninfo
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    cstart = up(batch.chrom_start)
    minus = up(np.array([s == '-' for s in batch.alignstrand], np.uint8))
    clen = up(np.array([chrom2len.get(c, -1) for c in batch.chrom], np.int64)) if chrom2len is not None else None
    rs = re_ = None
    if rg_chrom is not None:
        rlen = np.diff(batch.ev_off)
        same = np.array([c == rg_chrom for c in batch.chrom], bool)
        lo = batch.chrom_start if rg_start is None else np.full(n_reads, rg_start, np.int64)
        hi = batch.chrom_start + rlen if rg_end is None else np.full(n_reads, rg_end, np.int64)
        dead = ~same | (lo >= batch.chrom_start + rlen) | (hi <= batch.chrom_start)           # :307-308, :326-327
        rs, re_ = up(np.where(dead, 0, lo).astype(np.int64)), up(np.where(dead, 0, hi).astype(np.int64))
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)
    cap = max(1024, n_events // 8)
    text = "".join(motifs).encode("ascii")
    while True:
        i64 = dict(dtype=torch.int64, device=device)
        out = (torch.empty(cap, dtype=torch.int32, device=device), torch.empty(cap, **i64), torch.empty(cap, **i64),
               torch.empty(cap, **i64))
        n = C.c_int64(0)
        with torch.cuda.device(device):
            rc = L.dsp_find_sites(device.index, ptr(d["ev_base"]), C.c_void_p(d["ev_off"].data_ptr()), n_reads, n_events,
                                  text, len(motifs), mlen, int(methyloc), int(kmer_len), ptr(cstart), ptr(minus), ptr(clen),
                                  ptr(rs), ptr(re_), cap, *[ptr(t) for t in out], C.byref(n),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc == 4 and n.value > cap:                                # DSP_ERR_NOMEM: now the count is known
            cap = int(n.value)
            continue
        _native.check(rc, "dsp_find_sites")
        break
    k = int(n.value)
    site_read, site_ev, pos, pis = (t[:k] for t in out)
    sites = Sites(site_read.cpu().numpy(), site_ev.cpu().numpy(), pos.cpu().numpy(), pis.cpu().numpy())
    sites._dev = (device, site_read, site_ev)
    return sites


def sampleinfo(batch, sites):
    """The six leading columns of a feature / call_mods line (``call_modifications.py:312``)."""
    return ["\t".join([batch.chrom[r], str(int(p)), batch.alignstrand[r], str(int(q)), batch.readname[r], batch.strand[r]])
            for r, p, q in zip(sites.site_read, sites.pos, sites.pos_in_strand)]


def _packed_strings(xs):
    off = np.zeros(len(xs) + 1, np.int64)
    np.cumsum([len(x) for x in xs], out=off[1:])
    return np.frombuffer("".join(xs).encode("ascii"), np.uint8), off


def sampleinfo_packed(batch, sites, nthreads=None):
    """The same six columns packed back to back for ``dsp_format_calls``: (uint8 text, int64 offsets),
    written by ``dsp_format_sampleinfo`` (host threads)."""
    import os
    n = len(sites)
    if any(len(x) != 1 for x in batch.alignstrand) or any(len(x) != 1 for x in batch.strand):
        raise ValueError("alignstrand / strand must be single characters ('+'/'-', 't'/'c')")
    L = _native.lib()
    ctext, coff = _packed_strings(batch.chrom)
    ntext, noff = _packed_strings(batch.readname)
    astr = np.frombuffer("".join(batch.alignstrand).encode("ascii"), np.uint8)
    sstr = np.frombuffer("".join(batch.strand).encode("ascii"), np.uint8)
    per_read = (np.diff(coff) + np.diff(noff))[sites.site_read].sum() if n else 0
    cap = int(per_read) + n * 48
    text = np.empty(max(cap, 1), np.uint8)
    off = np.zeros(n + 1, np.int64)
    sr, pos, pis = (np.ascontiguousarray(a) for a in (sites.site_read, sites.pos, sites.pos_in_strand))
    p = lambda a: a.ctypes.data
    _native.check(L.dsp_format_sampleinfo(p(ctext), p(coff), p(ntext), p(noff), p(astr), p(sstr), p(sr), p(pos), p(pis), n,
                                          p(text), cap, p(off), int(nthreads or min(16, os.cpu_count() or 1))),
                  "dsp_format_sampleinfo")
    return text[:int(off[n])], off


def extract_tensors(batch, sites, kmer_len=13, signals_len=16, normalize_method="mad", round_stats=False,
                    drawn=None, seed=0, device=None, out=None, dtype=torch.float32):
    """Run ``dsp_extract_features``: -> dict of the five float32 CUDA tensors ``ModelBiLSTM.forward``
    takes (``kmer, base_means, base_stds, base_signal_lens, signals``) plus ``read_shift`` /
    ``read_scale`` (float64 per read: the median and MAD ``_normalize_signals`` used).
    ``drawn``: optional (n_sites, kmer_len, signals_len) int32 subsample offsets to replay (parity).
    ``out``: a dict returned by an earlier call with the same shapes, to be overwritten (no allocation).
    ``dtype=torch.float64`` (``dsp_extract_features_f64``) returns the values before the float32 narrowing,
    which is what the feature file prints."""
    if normalize_method not in ("mad", "zscore"):
        raise ValueError("")                                            # _normalize_signals, :184-185
    if not torch.cuda.is_available():
        raise _native.DspError("dsp_extract_features needs a CUDA device; there is no CPU fallback")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L = _native.lib()
    d = batch.to_device(device)
    n = len(sites)
    T, S = int(kmer_len), int(signals_len)
    site_read, site_ev = sites.to_device(device)
    drawn_t = None
    if drawn is not None:
        drawn_t = torch.as_tensor(np.ascontiguousarray(drawn, np.int32)).to(device)
        if tuple(drawn_t.shape) != (n, T, S):
            raise ValueError("drawn must have shape (n_sites, kmer_len, signals_len)")
    if dtype not in (torch.float32, torch.float64):
        raise ValueError("dtype must be torch.float32 or torch.float64")
    f32 = dict(dtype=dtype, device=device)
    if out is None:
        out = dict(kmer=torch.empty((n, T), **f32), base_means=torch.empty((n, T), **f32),
                   base_stds=torch.empty((n, T), **f32), base_signal_lens=torch.empty((n, T), **f32),
                   signals=torch.empty((n, T, S), **f32),
                   read_shift=torch.empty(batch.n_reads, dtype=torch.float64, device=device),
                   read_scale=torch.empty(batch.n_reads, dtype=torch.float64, device=device))
    elif (tuple(out["signals"].shape) != (n, T, S) or out["read_shift"].shape[0] != batch.n_reads
          or out["signals"].device != device or out["signals"].dtype != dtype):
        raise ValueError("out does not match this batch")
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)
    with torch.cuda.device(device):
        st = torch.cuda.current_stream().cuda_stream
        entry = L.dsp_extract_features if dtype == torch.float32 else L.dsp_extract_features_f64
        rc = entry(
            device.index, ptr(d["raw"]), C.c_void_p(d["raw_off"].data_ptr()), ptr(d["scaling"]), ptr(d["offset"]),
            batch.n_reads, ptr(d["ev_start"]), ptr(d["ev_len"]), ptr(d["ev_base"]), ptr(site_read), ptr(site_ev), n,
            T, S, 0 if normalize_method == "mad" else 1, 1 if round_stats else 0, ptr(drawn_t), int(seed) & (2 ** 64 - 1),
            C.c_void_p(out["read_shift"].data_ptr()), C.c_void_p(out["read_scale"].data_ptr()),
            ptr(out["kmer"]), ptr(out["base_means"]), ptr(out["base_stds"]), ptr(out["base_signal_lens"]),
            ptr(out["signals"]), C.c_void_p(st))
    _native.check(rc, "dsp_extract_features")
    return out


def format_features(batch, sites, means, stds, lens, signals, methy_label, nthreads=None, as_array=False):
    """bytes of the feature-file lines of the given sites (``_features_to_str``, ``:381-395``), written by
    ``dsp_format_features``.  means/stds/lens (n, T) and signals (n, T, S): float64 host arrays as
    ``extract_tensors(..., round_stats=True, dtype=torch.float64)`` returns them."""
    import os
    n = len(sites)
    if n == 0:
        return np.zeros(0, np.uint8) if as_array else b""
    L = _native.lib()
    arr = [np.ascontiguousarray(a, np.float64) for a in (means, stds, lens, signals)]
    T, S = arr[3].shape[1], arr[3].shape[2]
    nb = (T - 1) // 2
    letters = np.ascontiguousarray(batch.ev_base[sites.site_ev[:, None] + np.arange(-nb, nb + 1)[None, :]])
    info_text, info_off = sampleinfo_packed(batch, sites, nthreads)
    info_text = np.ascontiguousarray(info_text)
    cap = int(info_off[n]) + n * (T * (2 * 26 + 22 + S * 26) + 64)
    # sizes vary a lot (most values are short): ask first with a typical budget, grow on DSP_ERR_NOMEM
    budget = int(info_off[n]) + n * (T * (2 * 10 + 4 + S * 10) + 32)
    p = lambda a: a.ctypes.data
    used = C.c_int64(0)
    while True:
        out = np.empty(min(budget, cap), np.uint8)
        rc = L.dsp_format_features(p(info_text), p(info_off), p(letters), p(arr[0]), p(arr[1]), p(arr[2]), p(arr[3]),
                                   int(methy_label), n, T, S, p(out), out.shape[0], C.byref(used),
                                   int(nthreads or min(16, os.cpu_count() or 1)))
        if rc == 4 and used.value > out.shape[0]:
            budget = int(used.value)
            continue
        _native.check(rc, "dsp_format_features")
        return out[:used.value] if as_array else out[:used.value].tobytes()


def extract_to_file(args):
    """``deepsignal_plant extract`` (``extract_features.py:563-633``) for a decoded-reads archive: the
    feature file the reference writes (same lines, ``_features_to_str``), produced by the device kernels.
    Returns the number of sites written."""
    import gzip
    import os
    import time
    start = time.time()
    if not os.path.exists(args.fast5_dir):
        raise ValueError("--fast5_dir not set right!")
    if os.path.isdir(args.fast5_dir):
        raise ValueError("--fast5_dir is a directory of fast5 files: reading fast5 (h5py) is outside this implementation; "
                         "decode the reads once into an archive with extract_features.save_reads and pass the .npz")
    if str(getattr(args, "w_is_dir", "no")).lower() in ("yes", "true", "t", "1"):
        raise ValueError("--w_is_dir yes (one file per batch) is not supported; the features go to one file")
    allreads = load_reads(args.fast5_dir)
    motif_seqs = get_motif_seqs(args.motifs, str(args.is_dna).lower() in ("yes", "true", "t", "1"))
    chrom2len = get_contig2len(args.reference_path) if args.reference_path else None
    positions = _read_position_file(args.positions) if args.positions else None
    regioninfo = parse_region_str(args.region)
    dev = torch.device("cuda", 0)
    path = args.write_path + (".gz" if args.gzip and not args.write_path.endswith(".gz") else "")
    step = max(1, int(args.f5_batch_size))
    total = 0
    pinned = {}
    from . import feature_bin
    if args.write_path.endswith(feature_bin.SUFFIX):
        # binary hand-off to call_mods: the float32 the text file's values become in FloatTensor, no text in between
        if args.gzip:
            raise ValueError("--gzip does not apply to a binary feature file (%s)" % feature_bin.SUFFIX)
        nth = int(getattr(args, "host_threads", 0) or 0) or min(16, os.cpu_count() or 1)
        with feature_bin.FeatureBinWriter(args.write_path, args.seq_len, args.signal_len) as w:
            for lo in range(0, allreads.n_reads, step):
                batch = allreads.slice(lo, min(lo + step, allreads.n_reads))
                sites = find_sites_device(batch, motif_seqs, args.mod_loc, chrom2len, args.seq_len, positions, regioninfo, dev)
                if len(sites) == 0:
                    continue
                t = extract_tensors(batch, sites, args.seq_len, args.signal_len, args.normalize_method, True, seed=total, device=dev)
                info_text, info_off = sampleinfo_packed(batch, sites, nth)             # host work under the kernels
                host = [t[k].cpu().numpy() for k in ("kmer", "base_means", "base_stds", "base_signal_lens", "signals")]
                w.write(*host, args.methy_label, info_text, info_off)
                total += len(sites)
        print("[extract] {} sites from {} reads in {:.2f} seconds".format(total, allreads.n_reads, time.time() - start))
        return total
    with (gzip.open(path, "wb") if args.gzip else open(path, "wb")) as wf:
        for lo in range(0, allreads.n_reads, step):
            batch = allreads.slice(lo, min(lo + step, allreads.n_reads))
            sites = find_sites_device(batch, motif_seqs, args.mod_loc, chrom2len, args.seq_len, positions, regioninfo, dev)
            if len(sites) == 0:
                continue
            t = extract_tensors(batch, sites, args.seq_len, args.signal_len, args.normalize_method, True,
                                seed=total, device=dev, dtype=torch.float64)
            host = []
            for k in ("base_means", "base_stds", "base_signal_lens", "signals"):     # page-locked landing buffers, reused
                need = t[k].numel()
                if k not in pinned or pinned[k].numel() < need:
                    pinned[k] = torch.empty(int(need * 1.25), dtype=torch.float64).pin_memory()
                dst = pinned[k][:need].view(t[k].shape)
                dst.copy_(t[k], non_blocking=True)
                host.append(dst)
            torch.cuda.current_stream(dev).synchronize()
            text = format_features(batch, sites, *[h.numpy() for h in host], args.methy_label, nthreads=int(getattr(args, 'host_threads', 0) or 0) or min(32, os.cpu_count() or 1),
                                   as_array=True)
            wf.write(memoryview(text))
            total += len(sites)
    print("[extract] {} sites from {} reads in {:.2f} seconds".format(total, allreads.n_reads, time.time() - start))
    return total


def _extract_features(reads, normalize_method, motif_seqs, methyloc, chrom2len, kmer_len, signals_len,
                      methy_label, positions, regioninfo, drawn=None, seed=0, device=None):
    """The reference's ``_extract_features`` (``:280-378``) for decoded reads instead of fast5 paths:
    -> (features_list, error) with the same 12-tuples
    ``(chrom, pos, alignstrand, pos_in_strand, readname, strand, k_mer, signal_means, signal_stds,
    signal_lens, k_signals_rect, methy_label)``.  Values are what the device produced in float32
    (the precision every consumer -- ``FloatTensor`` -- reads them in)."""
    batch = pack_reads(reads)
    sites = find_sites(batch, motif_seqs, methyloc, chrom2len, kmer_len, positions, regioninfo)
    if len(sites) == 0:
        return [], 0
    t = extract_tensors(batch, sites, kmer_len, signals_len, normalize_method, False, drawn, seed, device)
    host = {k: t[k].cpu().numpy() for k in ("base_means", "base_stds", "base_signal_lens", "signals")}
    num_bases = (kmer_len - 1) // 2
    feats = []
    for i in range(len(sites)):
        r, ev = int(sites.site_read[i]), int(sites.site_ev[i])
        k_mer = bytes(batch.ev_base[ev - num_bases:ev + num_bases + 1]).decode()
        feats.append((batch.chrom[r], int(sites.pos[i]), batch.alignstrand[r], int(sites.pos_in_strand[i]),
                      batch.readname[r], batch.strand[r], k_mer, list(host["base_means"][i]), list(host["base_stds"][i]),
                      [int(x) for x in host["base_signal_lens"][i]], host["signals"][i].tolist(), methy_label))
    return feats, 0
