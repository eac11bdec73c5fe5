"""ORACLE (test infrastructure, not product code): restatement of the per-site output
contract of ``_call_mods`` (reference ``call_modifications.py:163,175-188``).

Given the softmax output of the model for one batch it reproduces the label choice and
the text of each call_mods line:

    chrom  pos  strand  pos_in_strand  readname  read_strand  prob_0  prob_1  label  5mer

Semantics that matter: probabilities are numpy *float32*
scalars; ``prob_0_norm = round(p0 / (p0 + p1), 6)`` and ``prob_1_norm = round(1 -
prob_0_norm, 6)`` are evaluated in float32; ``str()`` of a float32 prints the shortest
round-trip repr ('0.5', '1e-06', '0.0'); the label is the argmax of the *unrounded*
probabilities (``torch.max(vlogits.data, 1)``, ``:163``, first index on exact ties); the
5-mer is the centre window ``kmer[c-2:c+3]`` clipped to the k-mer (``:181-184``).

Pinned by ``tests/golden/callmods_*.tsv.gz`` produced by the reference's own
``_call_mods`` (``oracle/make_golden.py``).
"""
from __future__ import annotations

import numpy as np

# reference utils/process_utils.py:22-25 (base2code_dna inverted, :29)
CODE2BASE = "ACGTNWSMKRYBVDHZ"


def call_lines(sampleinfo, kmers, probs):
    """sampleinfo: list[str] (6 tab-joined columns); kmers: (N,T) int codes;
    probs: (N,2) float32 softmax. Returns (lines, labels)."""
    probs = np.asarray(probs, np.float32)
    labels = np.argmax(probs, axis=1)
    lines = []
    kmers = np.asarray(kmers).astype(np.int64)
    T = kmers.shape[1]
    c = T // 2
    lo, hi = max(c - 2, 0), min(c + 3, T)
    for n in range(probs.shape[0]):
        p0, p1 = probs[n, 0], probs[n, 1]            # numpy float32 scalars
        p0n = round(p0 / (p0 + p1), 6)
        p1n = round(1 - p0n, 6)
        five = "".join(CODE2BASE[b] for b in kmers[n, lo:hi])
        lines.append("\t".join((sampleinfo[n], str(p0n), str(p1n), str(labels[n]), five)))
    return lines, labels
