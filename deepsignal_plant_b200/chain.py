"""``call_mods`` -> ``call_freq`` without text in between (BASELINE.json configs[4]).

The reference couples its two commands through a TSV file: ``_call_mods`` prints
``prob_0_norm = round(p0 / (p0 + p1), 6)`` and ``prob_1_norm = round(1 - prob_0_norm, 6)`` as ``str(numpy.float32)``
(``call_modifications.py:177-188``) and ``ModRecord`` parses them back with ``float()`` (``utils/txt_formater.py:16-17``).
``records_from_probs`` computes, on the device, exactly the float64 values that round trip yields, so the calls of
a batch can stay in HBM as the record columns ``dsp_freq_aggregate`` / ``dsp_freq_aggregate_distributed`` take:

* ``numpy.round(x, 6)`` on float32 is ``rint(x * 1e6f) / 1e6f``: the printed value is ``k / 10^6`` for the integer
  ``k = rint(x * 1e6f)``;
* ``str(numpy.float32)`` prints the shortest decimal that round-trips; float32 spacing on [0, 1] is below 6e-8, so
  two different 6-decimal values never share a float32 and that shortest decimal is ``k / 10^6`` itself;
* ``float("0.dddddd")`` is the correctly rounded double of ``k / 10^6``, and so is the IEEE division
  ``double(k) / 1e6`` (both operands exact).

``DeviceCalls`` accumulates the record columns of successive batches in preallocated device buffers.
"""
from __future__ import annotations

import numpy as np


def records_from_probs(probs, labels=None):
    """probs: (n, 2) float32 CUDA tensor (what ``ModelBiLSTM.forward`` returns second); labels: (n,) int32 argmax or
    None.  -> (p0, p1, label): float64, float64, int32 device tensors = what ``ModRecord`` would parse from the
    line ``_call_mods`` prints for each site."""
    import torch
    p = probs.to(torch.float32)
    # every division has a TENSOR divisor: torch's CUDA kernels turn `tensor / python_scalar` into a multiplication by
    # the reciprocal, which is not the correctly rounded quotient numpy and float() compute
    m32 = torch.tensor(1e6, dtype=torch.float32, device=p.device)
    m64 = torch.tensor(1e6, dtype=torch.float64, device=p.device)
    x = p[:, 0] / (p[:, 0] + p[:, 1])                          # float32, call_modifications.py:177
    k0 = torch.round(x * m32)                                  # rint in float32: integer-valued, <= 1e6
    p0n = k0 / m32                                             # the float32 numpy.round returns
    k1 = torch.round((1.0 - p0n) * m32)                        # :178, float32 throughout
    if labels is None:
        labels = torch.max(p, 1)[1]                            # :163 (first index on ties)
    return k0.to(torch.float64) / m64, k1.to(torch.float64) / m64, labels.to(torch.int32)


class DeviceCalls:
    """Record columns (key, p0, p1, label) of up to ``capacity`` per-read calls, resident in HBM, appended batch by
    batch in call order (= the order ``call_mods`` would write its lines)."""

    def __init__(self, capacity, device):
        import torch
        self.dev = torch.device("cuda", device) if isinstance(device, int) else device
        self.key = torch.empty(capacity, dtype=torch.int64, device=self.dev)
        self.p0 = torch.empty(capacity, dtype=torch.float64, device=self.dev)
        self.p1 = torch.empty(capacity, dtype=torch.float64, device=self.dev)
        self.label = torch.empty(capacity, dtype=torch.int32, device=self.dev)
        self.n = 0

    def append(self, keys, probs, labels=None):
        m = int(keys.shape[0])
        if self.n + m > self.key.shape[0]:
            raise ValueError("DeviceCalls is full (%d + %d > %d)" % (self.n, m, self.key.shape[0]))
        p0, p1, lab = records_from_probs(probs, labels)
        s = slice(self.n, self.n + m)
        self.key[s], self.p0[s], self.p1[s], self.label[s] = keys, p0, p1, lab
        self.n += m

    def columns(self):
        return self.key[:self.n], self.p0[:self.n], self.p1[:self.n], self.label[:self.n]


def text_round_trip(probs):
    """HOST restatement used by the tests: the float64 pair ``ModRecord`` parses from what ``_call_mods`` prints."""
    from .call_modifications import normalise_probs
    p0n, p1n = normalise_probs(np.asarray(probs, np.float32))
    return (np.array([float(s) for s in p0n.astype(str)], np.float64),
            np.array([float(s) for s in p1n.astype(str)], np.float64))
