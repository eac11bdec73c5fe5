"""TEST stand-in for the device steps of the multi-rank call_freq (``freq_dist.DeviceBackend``), so that the
host logic -- byte shards, global record indices, chromosome tables, row order, slice writer -- can run under
world-size-2 ``gloo`` on a CPU box.  It has the backend's two calls with the backend's contract, computed with
the oracle's semantics (``oracle/freq_oracle.py``: float64 left-to-right per key) on all-gathered data.  Lives
in tests/: the package has no CPU path."""
import numpy as np

from deepsignal_plant_b200 import freq_dist as fd


class StandInBackend:
    def __init__(self, grp):
        self.grp = grp

    def aggregate(self, keys, p0, p1, label, gidx_base, bounds, prob_cf, balanced=True):
        parts = self.grp.all_gather_object((np.asarray(keys), np.asarray(p0), np.asarray(p1), np.asarray(label)))
        k = np.concatenate([x[0] for x in parts]); a = np.concatenate([x[1] for x in parts])
        b = np.concatenate([x[2] for x in parts]); lab = np.concatenate([x[3] for x in parts])
        table = {}
        for i in range(len(k)):                                   # global file order
            if abs(a[i] - b[i]) < prob_cf:
                continue
            r = table.get(int(k[i]))
            if r is None:
                r = table[int(k[i])] = [i, 0.0, 0.0, 0, 0]
            r[1] += float(a[i]); r[2] += float(b[i])
            r[3 if lab[i] == 1 else 4] += 1
        every = sorted((r[0], key, r) for key, r in table.items())           # dict insertion order = by first record
        if balanced:                                                          # equal slices of that order
            w, me = self.grp.world, self.grp.rank
            mine = every[len(every) * me // w: len(every) * (me + 1) // w]
        else:                                                                 # rows whose first record is in my shard
            lo, hi = int(bounds[self.grp.rank]), int(bounds[self.grp.rank + 1])
            mine = [e for e in every if lo <= e[0] < hi]
        rows = np.zeros(len(mine), fd.SITE_ROW)
        for j, (first, key, r) in enumerate(mine):
            rows[j] = (key, first, r[1], r[2], r[3], r[4], r[3] + r[4], 0)
        return rows

    def route_rows(self, rows, field, bounds):
        parts = self.grp.all_gather_object(rows)
        inner = np.asarray(bounds[1:self.grp.world], np.uint64)
        out = []
        for p in parts:                                           # source-rank order, source order kept
            dest = (p[field][:, None] >= inner[None, :]).sum(1) if len(inner) else np.zeros(len(p), np.int64)
            out.append(p[dest == self.grp.rank])
        return np.concatenate(out) if out else rows[:0]
