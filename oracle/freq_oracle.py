"""ORACLE (test infrastructure, not product code): restatement of deepsignal-plant's
``call_freq`` per-site aggregation.

Follows, relative to /root/reference/deepsignal_plant:

* ``utils/txt_formater.py:8-26``  -- record parsing; site key is (chrom, pos) only
  (``:12``, strand is NOT part of the key); a record is callable iff
  ``abs(p0 - p1) >= prob_cf`` in float64 (``:23-26``);
* ``call_mods_freq.py:29-74``     -- per key, in file order: ``prob_0 += p0``,
  ``prob_1 += p1`` (float64, left to right), ``coverage += 1``, ``met``/``unmet`` by
  label; strand / pos_in_strand / kmer come from the first *callable* record of the key
  (``:54-59``); dict insertion order = first callable appearance;
* ``call_mods_freq.py:77-122``    -- optional sort by ``(chrom str, pos int)`` (``:87-90``);
  TSV ``"%s\\t%d\\t%s\\t%d\\t%.3f\\t%.3f\\t%d\\t%d\\t%d\\t%.4f\\t%s"`` (``:112-118``) or
  bedMethyl with ``int(round(rmet*100 + 0.001, 0))`` (``:106-110``).

Pinned by ``tests/golden/freq_*`` produced by the reference's own functions
(``oracle/make_golden.py``).
"""
from __future__ import annotations


def aggregate(lines, prob_cf=0.0, contig_name=None):
    """lines: iterable of call_mods text lines. Returns an insertion-ordered dict
    (chrom, pos) -> [strand, pos_in_strand, kmer, sum_p0, sum_p1, met, unmet, coverage]."""
    table = {}
    for line in lines:
        w = line.strip().split("\t")
        chrom, pos = w[0], int(w[1])
        if contig_name is not None and chrom != contig_name:
            continue
        p0, p1 = float(w[6]), float(w[7])
        if abs(p0 - p1) < prob_cf:
            continue
        row = table.get((chrom, pos))
        if row is None:
            row = table[(chrom, pos)] = [w[2], int(w[3]), w[9], 0.0, 0.0, 0, 0, 0]
        row[3] += p0
        row[4] += p1
        if int(w[8]) == 1:
            row[5] += 1
        else:
            row[6] += 1
        row[7] += 1
    return table


def render(table, is_sort=False, is_bed=False):
    """Text of the frequency table exactly as ``write_sitekey2stats`` writes it."""
    keys = list(table.keys())
    if is_sort:
        keys.sort()            # (chrom as str, pos as int), call_mods_freq.py:88
    out = []
    for chrom, pos in keys:
        strand, pis, kmer, s0, s1, met, unmet, cov = table[(chrom, pos)]
        if cov <= 0:
            continue
        rmet = float(met) / cov
        if is_bed:
            out.append("\t".join([chrom, str(pos), str(pos + 1), ".", str(cov), strand,
                                  str(pos), str(pos + 1), "0,0,0", str(cov),
                                  str(int(round(rmet * 100 + 0.001, 0)))]) + "\n")
        else:
            out.append("%s\t%d\t%s\t%d\t%.3f\t%.3f\t%d\t%d\t%d\t%.4f\t%s\n" % (
                chrom, pos, strand, pis, s0, s1, met, unmet, cov, rmet, kmer))
    return "".join(out)


def render_by_contig(lines, contigs, prob_cf=0.0, is_sort=False, is_bed=False):
    """Per-contig mode (``call_mods_freq.py:154-215,262-295``): one aggregation per contig, results
    concatenated in sorted temp-file-name order (``<result>.<contig>.<uuid>``, i.e. by contig + ".");
    contigs without records produce nothing."""
    lines = list(lines)
    out = []
    for contig in sorted(set(contigs), key=lambda c: c + "."):
        table = aggregate(lines, prob_cf, contig)
        if table:
            out.append(render(table, is_sort, is_bed))
    return "".join(out)
