"""Drop-in for the batch inference step of the reference's ``call_mods``.

Mirrors ``deepsignal_plant/call_modifications.py:130-192`` (``_call_mods``): same argument
tuple, same return triple ``(pred_str, accuracy, batch_num)``, same text per site

    chrom  pos  strand  pos_in_strand  readname  read_strand  prob_0  prob_1  label  5mer

but the per-site Python work of the reference (list -> tensor conversion ``:159-162``,
per-site renormalise/round/str-join loop ``:175-188``) is done with whole-batch array
operations, and the model call goes to the CUDA kernels of ``libdsp_b200``.
"""
from __future__ import annotations

import gzip
import os
import queue
import threading
import time

import numpy as np
import torch

from .models import ModelBiLSTM  # noqa: F401  (re-export, as the reference module does)

# reference utils/process_utils.py:22-29
base2code_dna = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'N': 4, 'W': 5, 'S': 6, 'M': 7, 'K': 8, 'R': 9,
                 'Y': 10, 'B': 11, 'V': 12, 'D': 13, 'H': 14, 'Z': 15}
code2base_dna = dict((v, k) for k, v in base2code_dna.items())
_CODE_LUT = np.frombuffer(b"ACGTNWSMKRYBVDHZ", dtype="S1")


def str2bool(v):
    """``utils/process_utils.py:54-56``."""
    return v.lower() in ("yes", "true", "t", "1")


def FloatTensor(tensor, device=0):
    """``utils/constants_torch.py:10-13``: nested lists -> float32 tensor on cuda:device.
    Goes through one numpy conversion and a pinned staging buffer instead of
    ``torch.tensor(list)``; raises without a GPU (no CPU path)."""
    if not torch.cuda.is_available():
        raise RuntimeError("deepsignal_plant_b200 needs a CUDA device; there is no CPU path")
    a = torch.from_numpy(np.ascontiguousarray(np.asarray(tensor, dtype=np.float32)))
    return a.pin_memory().to("cuda:{}".format(device), non_blocking=True)


def normalise_probs(probs):
    """float32 arithmetic of ``call_modifications.py:177-179``:
    ``p0n = round(p0/(p0+p1), 6)``, ``p1n = round(1 - p0n, 6)`` evaluated in float32."""
    probs = np.asarray(probs, dtype=np.float32)
    p0, p1 = probs[:, 0], probs[:, 1]
    p0n = np.round(p0 / (p0 + p1), 6)
    p1n = np.round(np.float32(1) - p0n, 6)
    return p0n.astype(np.float32), p1n.astype(np.float32)


def kmer_centre(kmers, width=5):
    """Centre window of each k-mer as text (``:181-184``): kmer[c-2:c+3], clipped."""
    kmers = np.asarray(kmers)
    if kmers.dtype.kind == "f":
        kmers = kmers.astype(np.int64)
    T = kmers.shape[1]
    c = T // 2
    lo, hi = max(c - 2, 0), min(c + 3, T)
    w = hi - lo
    letters = np.ascontiguousarray(_CODE_LUT[kmers[:, lo:hi]])
    return letters.view("S%d" % w).reshape(-1).astype("U%d" % w)


def format_calls(sampleinfo, kmers, probs, labels):
    """Text lines of one batch, identical to the reference's per-site loop (``:175-188``)."""
    p0n, p1n = normalise_probs(probs)
    s0 = p0n.astype(str)           # shortest float32 repr, same as str(np.float32)
    s1 = p1n.astype(str)
    lab = np.asarray(labels).astype(np.int64).astype(str)
    five = kmer_centre(kmers)
    return ["\t".join(t) for t in zip(sampleinfo, s0.tolist(), s1.tolist(), lab.tolist(), five.tolist())]


def _call_mods(features_batch, model, batch_size, device=0):
    """call modification from a batch of features (``call_modifications.py:130-192``).

    features_batch = (sampleinfo, kmers, base_means, base_stds, base_signal_lens, k_signals,
    labels) as Python lists (what ``_read_features_file`` produces) or numpy arrays.
    Returns (pred_str, accuracy, batch_num)."""
    sampleinfo, kmers, base_means, base_stds, base_signal_lens, k_signals, labels = features_batch
    n = len(sampleinfo)
    labels = np.reshape(labels, (len(labels)))
    kmers_a = np.asarray(kmers, dtype=np.float32).reshape(n, -1)
    means_a = np.asarray(base_means, dtype=np.float32).reshape(n, -1)
    stds_a = np.asarray(base_stds, dtype=np.float32).reshape(n, -1)
    lens_a = np.asarray(base_signal_lens, dtype=np.float32).reshape(n, -1)
    sig_a = np.asarray(k_signals, dtype=np.float32)
    dev = "cuda:{}".format(device)

    pred_str = []
    accuracys = []
    batch_num = 0
    pending = []
    for s in range(0, n, batch_size):
        e = min(s + batch_size, n)
        if e <= s:
            continue
        up = [torch.from_numpy(a[s:e]).to(dev, non_blocking=True) for a in (kmers_a, means_a, stds_a, lens_a, sig_a)]
        _, vprobs = model(*up)
        vlabels = getattr(model, "last_labels", None)
        if vlabels is None:
            vlabels = torch.max(vprobs.data, 1)[1]
        pending.append((s, e, vprobs, vlabels))
        batch_num += 1
    for s, e, vprobs, vlabels in pending:
        probs = vprobs.cpu().numpy()
        predicted = vlabels.cpu().numpy()
        accuracys.append(float(np.mean(labels[s:e] == predicted)))
        pred_str.extend(format_calls(sampleinfo[s:e], kmers_a[s:e], probs, predicted))
    accuracy = np.mean(accuracys) if len(accuracys) > 0 else 0
    return pred_str, accuracy, batch_num


# ---- queue-based workers with the reference's names and protocol -----------------------------------
# (call_modifications.py:55-127, 195-282).  `call_mods` below does not use them -- it streams pinned
# batches -- but code written against the reference's worker functions keeps working: same arguments,
# same queue items (7-tuples of Python lists, "kill" sentinel), same batch boundaries.

class SimpleQueue:
    """Minimal in-process queue with the methods the workers use (put/get/empty/qsize)."""

    def __init__(self):
        self._q = queue.Queue()

    def put(self, x):
        self._q.put(x)

    def get(self):
        return self._q.get()

    def empty(self):
        return self._q.empty()

    def qsize(self):
        return self._q.qsize()


def _read_features_file(features_file, features_batch_q, f5_batch_size=10):
    """``call_modifications.py:55-127``: parse the feature file (natively) and put batches of
    ``f5_batch_size`` reads -- cut exactly where the reference cuts them, when the read id in column 5
    has changed ``f5_batch_size`` times -- on the queue, then ``"kill"``."""
    from . import feature_io
    print("read_features process-{} starts".format(os.getpid()))
    r_num, b_num = 0, 0
    cur = [[], [], [], [], [], [], []]
    readid_pre = None
    reader = feature_io.FeatureFileReader(features_file, batch_sites=16384, pinned=False, slots=2,
                                          seq_len=None, signal_len=None)
    for blk in reader:
        cols = blk.as_reference_lists()
        for i, info in enumerate(cols[0]):
            readid = info.split("\t")[4]
            if readid_pre is not None and readid != readid_pre:
                r_num += 1
                if r_num % f5_batch_size == 0:
                    features_batch_q.put(tuple(cur))
                    cur = [[], [], [], [], [], [], []]
                    b_num += 1
            readid_pre = readid
            for c, col in zip(cur, cols):
                c.append(col[i])
    r_num += 1
    if len(cur[0]) > 0:
        features_batch_q.put(tuple(cur))
        b_num += 1
    features_batch_q.put("kill")
    print("read_features process-{} ending, read {} reads in {} f5-batches({})".format(os.getpid(), r_num, b_num,
                                                                                       f5_batch_size))


def _call_mods_q(model_path, features_batch_q, pred_str_q, success_file, args, device=0):
    """``call_modifications.py:195-259``: build the model from ``args``, load the checkpoint, then call
    every batch taken from ``features_batch_q`` until the ``"kill"`` sentinel (which is put back for
    sibling workers, ``:241-245``)."""
    print('call_mods process-{} starts'.format(os.getpid()))
    args.model_path = model_path
    model = load_model(args, device)
    batch_num_total = 0
    while True:
        features_batch = features_batch_q.get()
        if isinstance(features_batch, str) and features_batch == "kill":
            features_batch_q.put("kill")
            break
        pred_str, accuracy, batch_num = _call_mods(features_batch, model, args.batch_size, device)
        pred_str_q.put(pred_str)
        batch_num_total += batch_num
    print('call_mods process-{} ending, proceed {} feature-batches({})'.format(os.getpid(), batch_num_total,
                                                                               args.batch_size))


def _write_predstr_to_file(write_fp, predstr_q, is_gzip):
    """``call_modifications.py:262-282``."""
    print('write_process-{} starts'.format(os.getpid()))
    if is_gzip and not write_fp.endswith(".gz"):
        write_fp += ".gz"
    with (gzip.open(write_fp, "wt") if is_gzip else open(write_fp, "w")) as wf:
        while True:
            pred_str = predstr_q.get()
            if isinstance(pred_str, str) and pred_str == "kill":
                break
            wf.write("".join(line + "\n" for line in pred_str))
            wf.flush()
    print('write_process-{} finished'.format(os.getpid()))


# ---- the call_mods pipeline over a feature file (call_modifications.py:195-282, 532-640) -----------

def load_model(args, device=0):
    """Model construction and checkpoint loading exactly as ``_call_mods_q`` does it
    (``call_modifications.py:214-228``)."""
    model = ModelBiLSTM(args.seq_len, args.signal_len, args.layernum1, args.layernum2, args.class_num,
                        args.dropout_rate, args.hid_rnn,
                        args.n_vocab, args.n_embed, str2bool(args.is_base), str2bool(args.is_signallen),
                        module=args.model_type, device=device,
                        max_batch=getattr(args, "max_batch", 65536))
    para_dict = torch.load(args.model_path, map_location=torch.device("cpu"))
    model_dict = model.state_dict()
    model_dict.update(para_dict)
    model.load_state_dict(model_dict)
    del model_dict
    if not torch.cuda.is_available():
        raise RuntimeError("deepsignal_plant_b200 call_mods needs a CUDA device; there is no CPU path")
    model = model.cuda(device)
    model.eval()
    return model


def call_mods_stream(model, batches, write, depth=2, format_threads=None):
    """Drive the model over an iterable of ``feature_io.FeatureBatch`` (page-locked tensors):
    batch i+1 is submitted before batch i is collected, so its host->device copies overlap
    batch i's kernels; a formatter thread turns finished batches into output lines (native code on the host
    threads, outside the GIL) while this thread already waits for the next batch, and ``write(uint8 array)``
    receives each batch's lines in order.  A batch is in use until its lines are written: readers that recycle
    their slots need ``depth + 6`` of them (1 being filled, 2 queued in front of this function, ``depth`` in
    flight, 1 queued for the formatter, 1 being formatted, 1 spare).
    Returns (sites, mean per-batch accuracy against the label column, batches) -- the
    bookkeeping ``_call_mods_q`` keeps (``:230,255-257``)."""
    from . import feature_io
    C_ = model.num_classes
    if C_ != 2:
        raise ValueError("call_mods output format is defined for class_num == 2 (prob_0, prob_1)")
    outs = []
    inflight = []
    sites, acc, nb = 0, [], 0
    prof = {"wait for reader": 0.0, "submit": 0.0, "wait for device": 0.0, "wait for formatter": 0.0}
    fprof = {"format": 0.0, "hand to writer": 0.0}
    tick = time.perf_counter
    pin = torch.cuda.is_available()                  # result staging only; the model itself refuses to run without a GPU
    fq = queue.Queue(maxsize=1)
    ferr = []
    t_begin = tick()

    def formatter():                                 # the per-site text loop of _call_mods (:175-188), off the submit path
        while True:
            item = fq.get()
            if item is None:
                return
            if ferr:
                continue                             # keep draining so that the producer never blocks
            try:
                b, slot = item
                probs, labels = slot[1][:b.n].numpy(), slot[2][:b.n].numpy()
                t1 = tick()
                text = feature_io.format_calls(b, probs, labels, format_threads, as_array=True)     # no copy: the writer owns it
                t2 = tick()
                write(text)
                fprof["format"] += t2 - t1
                fprof["hand to writer"] += tick() - t2
                acc.append(float(np.mean(b.labels.numpy() == labels)))
                outs.append(slot)
            except BaseException as e:               # surfaced in the calling thread
                ferr.append(e)

    ft = threading.Thread(target=formatter, daemon=True)
    ft.start()

    def collect():
        nonlocal sites, nb
        b, tk, slot = inflight.pop(0)
        t0 = tick()
        model.wait_host(tk)
        t1 = tick()
        fq.put((b, slot))
        prof["wait for device"] += t1 - t0
        prof["wait for formatter"] += tick() - t1
        sites += b.n
        nb += 1

    try:
        it = iter(batches)
        while not ferr:
            t0 = tick()
            b = next(it, None)
            prof["wait for reader"] += tick() - t0
            if b is None:
                break
            cap = b.kmer.shape[0]
            slot = None
            for i in range(len(outs)):                   # the formatter thread only appends; this thread alone removes
                if outs[i][0].shape[0] >= cap:
                    slot = outs.pop(i)
                    break
            if slot is None:
                mk = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory() if pin else torch.empty(shape, dtype=dt)
                slot = (mk((cap, C_), torch.float32), mk((cap, C_), torch.float32), mk((cap,), torch.int32))
            lg, pr, lb = slot
            t0 = tick()
            tk = model.submit_host(*b.arrays(), lg[:b.n], pr[:b.n], lb[:b.n])
            prof["submit"] += tick() - t0
            inflight.append((b, tk, slot))
            if len(inflight) >= depth:
                collect()
        while inflight:
            collect()
    finally:
        fq.put(None)
        ft.join()
    if ferr:
        raise ferr[0]
    if os.environ.get("DSP_B200_PROFILE"):
        print("call_mods_stream: %d sites in %.3f seconds = %.0f sites/s (first batch in to last batch out)" % (sites, tick() - t_begin, sites / max(tick() - t_begin, 1e-9)))
        print("call_mods_stream host seconds: submitting thread: " + ", ".join("%s %.3f" % kv for kv in prof.items())
              + "; formatter thread: " + ", ".join("%s %.3f" % kv for kv in fprof.items()))
    return sites, (float(np.mean(acc)) if acc else 0.0), nb


def host_threads(args, world=1):
    """--host_threads: threads of the native parsers / formatters; 0 or absent = all cores, divided between the ranks
    (--nproc keeps the reference's meaning and default -- a number of worker processes -- and is not used for this)."""
    n = int(getattr(args, "host_threads", 0) or 0)
    if n <= 0:
        n = max(1, min(32, (os.cpu_count() or 1) // max(world, 1)))
    return n


def _shard_of_file(path, rank, world):
    """Contiguous byte shard of rank ``rank`` (lines that start inside it); ``None`` = whole file."""
    if world <= 1:
        return None
    size = os.path.getsize(path)
    return (size * rank // world, size * (rank + 1) // world)


class _CallBatch:
    """What ``feature_io.format_calls`` needs of a batch that never was text."""
    __slots__ = ("n", "kmer", "info_text", "info_off", "seq_len")

    def __init__(self, n, kmer, info_text, info_off, seq_len):
        self.n, self.kmer, self.info_text, self.info_off, self.seq_len = n, kmer, info_text, info_off, seq_len


def call_mods_from_reads(args, model, write, device=0):
    """The fast5 branch of ``call_mods`` (``call_modifications.py:560-577``:
    ``_read_features_from_fast5s`` -> ``_call_mods``, ``:285-323,361-443``) for reads that are already
    decoded (``extract_features.save_reads`` archive; h5py is absent here): per chunk of reads, site
    search on the host, ``dsp_extract_features`` -> ``dsp_forward`` on device tensors (features never
    become text and never leave the GPU), then the usual output lines.  Returns (sites, chunks)."""
    from . import extract_features as ef
    from . import feature_io
    prof = {"load": 0.0, "find_sites": 0.0, "launch": 0.0, "sampleinfo": 0.0, "format": 0.0}
    tick = time.perf_counter
    t0 = tick()
    allreads = ef.load_reads(args.input_path)
    prof["load"] = tick() - t0
    motif_seqs = ef.get_motif_seqs(args.motifs, str2bool(getattr(args, "is_dna", "yes")))
    chrom2len = ef.get_contig2len(args.reference_path) if getattr(args, "reference_path", None) else None
    positions = ef._read_position_file(args.positions) if getattr(args, "positions", None) else None
    regioninfo = ef.parse_region_str(getattr(args, "region", None))
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lo_r, hi_r = allreads.n_reads * rank // world, allreads.n_reads * (rank + 1) // world      # contiguous read shard
    dev = torch.device("cuda", device)
    max_batch = int(getattr(args, "max_batch", 65536))
    step = max(1, int(getattr(args, "f5_batch_size", 30)))
    sites_total, chunks = 0, 0
    pending = None

    def flush(item):
        # collecting chunk i-1 after chunk i is enqueued: its device->host reads do not stall the launches
        b, probs, labels = item
        t1 = tick()
        b.kmer = b.kmer.cpu().numpy()
        write(feature_io.format_calls(b, probs.cpu().numpy(), labels.cpu().numpy()))
        prof["format"] += tick() - t1

    # Upload, site search and extraction of chunk i+1 run on a side stream: the search has to tell the host how
    # many sites it found, and on the classifier's stream that read-back would wait for chunk i's whole forward.
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)
    lo = lo_r
    while lo < hi_r:
        hi = min(lo + step, hi_r)
        batch = allreads.slice(lo, hi)
        lo = hi
        t1 = tick()
        with torch.cuda.stream(side):
            sites = ef.find_sites_device(batch, motif_seqs, args.mod_loc, chrom2len, args.seq_len, positions, regioninfo, dev)
        prof["find_sites"] += tick() - t1
        n = len(sites)
        # --f5_batch_size only seeds the chunking: aim at one full model batch of sites per chunk
        step = max(1, min(int(step * max_batch / max(n, 1)), 4 * step + 64))
        if n == 0:
            continue
        t1 = tick()
        with torch.cuda.stream(side):
            t = ef.extract_tensors(batch, sites, args.seq_len, args.signal_len, args.normalize_method, False,
                                   seed=sites_total, device=dev)
        main.wait_stream(side)
        for v in t.values():
            v.record_stream(main)
        outs = []
        for a in range(0, n, max_batch):                       # the model's workspace holds max_batch sites
            z = min(a + max_batch, n)
            with torch.no_grad():
                _, probs = model(t["kmer"][a:z], t["base_means"][a:z], t["base_stds"][a:z],
                                 t["base_signal_lens"][a:z], t["signals"][a:z])
            outs.append((probs, model.last_labels))
        prof["launch"] += tick() - t1
        t1 = tick()
        info_text, info_off = ef.sampleinfo_packed(batch, sites)          # host work overlaps the kernels above
        prof["sampleinfo"] += tick() - t1
        if pending is not None:
            flush(pending)
        probs = outs[0][0] if len(outs) == 1 else torch.cat([o[0] for o in outs])
        labels = outs[0][1] if len(outs) == 1 else torch.cat([o[1] for o in outs])
        pending = (_CallBatch(n, t["kmer"], info_text, info_off, args.seq_len), probs, labels)
        sites_total += n
        chunks += 1
    if pending is not None:
        flush(pending)
    if os.environ.get("DSP_B200_PROFILE"):
        print("call_mods_from_reads host seconds: " + ", ".join("%s %.3f" % kv for kv in prof.items()))
    return sites_total, chunks


def call_mods(args):
    """``call_modifications.py:532-640`` for a feature-file input: read -> call -> write, output
    lines in file order.  Under ``torchrun`` (one process per GPU, RANK/WORLD_SIZE set) every rank
    takes a contiguous byte shard of the (uncompressed) feature file, writes its own part and rank
    0 concatenates the parts in rank order -- no collective touches the data path."""
    from . import feature_io
    print("[main] call_mods starts..")
    start = time.time()
    print("cuda availability: {}".format(torch.cuda.is_available()))
    model_path = os.path.abspath(args.model_path)
    if not os.path.exists(model_path):
        raise ValueError("--model_path is not set right!")
    input_path = os.path.abspath(args.input_path)
    if not os.path.exists(input_path):
        raise ValueError("--input_path does not exist!")
    if os.path.isdir(input_path):
        raise ValueError("--input_path is a directory of fast5 files: reading fast5 (h5py) is outside this "
                         "implementation; pass a feature file from `deepsignal_plant extract`, or decode the reads "
                         "once into an archive with extract_features.save_reads and pass the .npz")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = 0
    if world > 1:
        import torch.distributed as dist
        ndev = max(torch.cuda.device_count(), 1)
        device = int(os.environ.get("LOCAL_RANK", "0")) % ndev        # one GPU per rank; ranks may share one (tests)
        if not dist.is_initialized():
            if ndev >= world:
                dist.init_process_group("nccl", device_id=torch.device("cuda", device))
            else:
                dist.init_process_group("gloo")
        torch.cuda.set_device(device)
    args.model_path = model_path
    from_reads = input_path.endswith(".npz")                  # decoded reads instead of a feature file
    from . import feature_bin
    from_bin = not from_reads and feature_bin.is_feature_bin(input_path)      # binary feature hand-off (.dspf) instead of text
    gz_single = world > 1 and not from_reads and not from_bin and input_path.endswith(".gz")
    if gz_single and rank == 0:
        print("call_mods: a gzip feature file cannot be cut into byte shards; rank 0 processes it alone "
              "(decompress it to use all %d GPUs)" % world)
    # The reader starts before the model is built: its page-locked slots and the first batches are ready by the time the
    # weights are packed (_read_features_file, :55-127, is a process of its own in the reference too)
    rq = queue.Queue(maxsize=2)
    err = []
    rt = None
    depth = max(1, min(int(getattr(args, "stream_depth", 3) or 3), 6))           # batches in flight on the device
    if not from_reads:
        idle = gz_single and rank > 0                          # nothing to read on this rank
        if from_bin:
            # every rank reads a contiguous site range; sections go straight from the page cache into page-locked slots
            reader = feature_bin.FeatureBinReader(input_path, args.seq_len, args.signal_len, batch_sites=getattr(args, "max_batch", 65536),
                                                  slots=depth + 6, nthreads=int(getattr(args, "reader_threads", 0) or 0) or min(8, host_threads(args, world)))
            if world > 1:
                reader.site_range = (reader.total_sites * rank // world, reader.total_sites * (rank + 1) // world)
        else:
            reader = None if idle else feature_io.FeatureFileReader(
                input_path, args.seq_len, args.signal_len, batch_sites=getattr(args, "max_batch", 65536), slots=depth + 6,
                nthreads=int(getattr(args, "reader_threads", 0) or 0) or host_threads(args, world), byte_range=None if gz_single else _shard_of_file(input_path, rank, world))

        stop_reading = threading.Event()

        def read():
            try:
                if torch.cuda.is_available():
                    torch.cuda.set_device(device)              # page-locked slots belong to this rank's GPU context
                for b in (reader or ()):
                    if stop_reading.is_set():
                        break
                    rq.put(b)
            except BaseException as e:                         # surfaced in the main thread
                err.append(e)
            rq.put(None)

        rt = threading.Thread(target=read, daemon=True)
        rt.start()

        def abandon_reader():                                  # the model could not be built: let the reader thread end
            stop_reading.set()
            while rt.is_alive():
                try:
                    rq.get(timeout=0.1)
                except queue.Empty:
                    pass
    try:
        model = load_model(args, device)
    except BaseException:
        if rt is not None:
            abandon_reader()
        raise
    args.input_path = input_path

    result_file = args.result_file
    if args.gzip and not result_file.endswith(".gz"):
        result_file += ".gz"                                   # call_modifications.py:264-266
    my_file = result_file if world == 1 else "%s.part%05d" % (result_file, rank)
    opener = (lambda p: gzip.open(p, "wb")) if args.gzip else (lambda p: open(p, "wb"))
    wq = queue.Queue(maxsize=8)

    werr = []

    def writer():                                              # _write_predstr_to_file (:262-282)
        # a failing write (disk full, bad path) must not leave the producers blocked on a full queue: remember the
        # error, keep draining until the sentinel, re-raise in the main thread after the join
        wf = None
        try:
            wf = opener(my_file)
        except BaseException as e:
            werr.append(e)
        while True:
            item = wq.get()
            if item is None:
                break
            if wf is not None and not werr:
                try:
                    wf.write(item)
                except BaseException as e:
                    werr.append(e)
        if wf is not None:
            try:
                wf.close()
            except BaseException as e:
                werr.append(e)

    wt = threading.Thread(target=writer, daemon=True)
    wt.start()

    # --freq_out: the calls of this run also feed call_freq directly -- every batch's lines are parsed back into record
    # columns while they are still in memory (dsp_parse_calls, the parser call_freq uses on files), the table is
    # aggregated at the end; under torchrun the ranks' shards meet through the NVLink exchange (freq_dist.py)
    freq_out = getattr(args, "freq_out", None)
    freq_parts = []

    def sink(data):
        wq.put(data)
        if freq_out:
            from . import call_mods_freq as cf
            r = cf.parse_calls_buffer(data, nthreads=host_threads(args, world))
            if r is None:
                r = cf.parse_lines(bytes(data).decode().splitlines())
            freq_parts.append(r)

    def finish_writer():
        wq.put(None)
        wt.join()
        if werr:
            raise werr[0]

    def finish_freq():
        if not freq_out:
            return
        from . import call_mods_freq as cf
        rec = cf.Records.concat(freq_parts)
        prob_cf, is_sort, is_bed = getattr(args, "freq_prob_cf", 0.5), getattr(args, "freq_sort", False), getattr(args, "freq_bed", False)
        if world > 1:
            from . import freq_dist
            table, _, _ = freq_dist.call_freq_distributed(None, prob_cf, freq_out, is_sort, is_bed, False, device=device, records=rec)
        else:
            table = cf.aggregate_records(rec, prob_cf, device=device)
            cf.write_sitekey2stats(table, freq_out, is_sort, is_bed, False)
        if rank == 0:
            print("call_freq of this run: {} of {} calls used -> {}".format(table.n_used, table.n_records, freq_out))
    if from_reads:
        sites, nb = call_mods_from_reads(args, model, sink, device)
        finish_writer()
        print("call_mods rank {}: {} sites in {} read-batches".format(rank, sites, nb))
        _merge_parts(result_file, rank, world)
        finish_freq()
        print("[main] call_mods costs %.2f seconds.." % (time.time() - start))
        return sites

    def batches():
        while True:
            b = rq.get()
            if b is None:
                return
            yield b

    sites, accuracy, nb = call_mods_stream(model, batches(), sink, depth, int(getattr(args, "format_threads", 0) or 0) or min(8, host_threads(args, world)))
    rt.join()
    finish_writer()
    if err:
        raise err[0]
    print("call_mods rank {}: {} sites in {} feature-batches".format(rank, sites, nb))
    _merge_parts(result_file, rank, world)
    finish_freq()
    print("[main] call_mods costs %.2f seconds.." % (time.time() - start))
    return sites


def _copy_into(src_path, dst_path, offset):
    """Copy the whole of ``src_path`` into ``dst_path`` at byte ``offset`` (in the kernel where the filesystem allows)."""
    size = os.path.getsize(src_path)
    with open(src_path, "rb") as src, open(dst_path, "r+b") as dst:
        done = 0
        if hasattr(os, "copy_file_range"):
            try:
                while done < size:
                    k = os.copy_file_range(src.fileno(), dst.fileno(), min(size - done, 1 << 30), done, offset + done)
                    if k <= 0:
                        break
                    done += k
            except OSError:                                     # across filesystems / not supported: plain copy below
                pass
        src.seek(done)
        dst.seek(offset + done)
        while done < size:
            blk = src.read(min(1 << 24, size - done))
            if not blk:
                raise IOError("%s shrank while it was merged" % src_path)
            dst.write(blk)
            done += len(blk)


def _merge_parts(result_file, rank, world):
    """The per-rank parts become one file in rank order (contiguous shards -> input order): the ranks exchange their
    part sizes, rank 0 sizes the result, and every rank copies ITS part to its offset -- the merge is as parallel as
    the run was, instead of one process re-writing everybody's output (58 GB of calls for 10^9 sites)."""
    if world <= 1:
        return
    import torch.distributed as dist
    part = "%s.part%05d" % (result_file, rank)
    sizes = [None] * world
    dist.all_gather_object(sizes, os.path.getsize(part))
    if rank == 0:
        with open(result_file, "wb") as out:
            out.truncate(sum(sizes))
    dist.barrier()
    _copy_into(part, result_file, sum(sizes[:rank]))
    os.remove(part)
    dist.barrier()
