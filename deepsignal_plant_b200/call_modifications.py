"""Drop-in for the batch inference step of the reference's ``call_mods``.

Mirrors ``deepsignal_plant/call_modifications.py:130-192`` (``_call_mods``): same argument
tuple, same return triple ``(pred_str, accuracy, batch_num)``, same text per site

    chrom  pos  strand  pos_in_strand  readname  read_strand  prob_0  prob_1  label  5mer

but the per-site Python work of the reference (list -> tensor conversion ``:159-162``,
per-site renormalise/round/str-join loop ``:175-188``) is done with whole-batch array
operations, and the model call goes to the CUDA kernels of ``libdsp_b200``.
"""
from __future__ import annotations

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
