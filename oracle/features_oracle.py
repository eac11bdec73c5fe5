"""ORACLE (test infrastructure, not product code): per-line restatement of the reference's
feature-file reader.

Follows ``deepsignal_plant/call_modifications.py:55-127`` (``_read_features_file``): every line is
``strip()``-ed and split on tabs; columns 0-5 are re-joined as the sample info (``:89``), column 6
goes through ``base2code_dna`` (``utils/process_utils.py:22-29``), columns 7/8 are ``float``
lists, 9 an ``int`` list, 10 ``;``-separated groups of ``float`` lists, 11 an ``int`` label
(``:90-95``).  Batching by read id / ``f5_batch_size`` (``:97-113``) only decides where the
stream is cut and is restated in ``batch_sizes``.  Pinned by ``tests/golden/features_small*``
(``oracle/make_golden_features.py`` runs the reference reader itself)."""
from __future__ import annotations

base2code_dna = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'N': 4, 'W': 5, 'S': 6, 'M': 7, 'K': 8, 'R': 9,
                 'Y': 10, 'B': 11, 'V': 12, 'D': 13, 'H': 14, 'Z': 15}


def parse_line(line):
    words = line.strip().split("\t")
    return ("\t".join(words[0:6]), [base2code_dna[x] for x in words[6]], [float(x) for x in words[7].split(",")],
            [float(x) for x in words[8].split(",")], [int(x) for x in words[9].split(",")],
            [[float(y) for y in x.split(",")] for x in words[10].split(";")], int(words[11]))


def read_features(lines):
    """-> 7 parallel lists, like one queue item of the reference."""
    cols = ([], [], [], [], [], [], [])
    for line in lines:
        for c, v in zip(cols, parse_line(line)):
            c.append(v)
    return cols


def batch_sizes(lines, f5_batch_size):
    """Sizes of the batches the reference puts on its queue (cut when the read id in column 4 has
    changed ``f5_batch_size`` times, ``:97-113``)."""
    sizes, cur, r_num, prev = [], 0, 0, None
    for line in lines:
        rid = line.strip().split("\t")[4]
        if prev is not None and rid != prev:
            r_num += 1
            if r_num % f5_batch_size == 0:
                sizes.append(cur)
                cur = 0
        prev = rid
        cur += 1
    if cur:
        sizes.append(cur)
    return sizes
