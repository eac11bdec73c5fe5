#!/usr/bin/env python
"""Instruction counts per kernel, from `cuobjdump -sass`, for the mnemonics that identify the Blackwell paths
(runs in the build container: no GPU needed).

    python tools/sass_mnemonics.py [object or .so ...] > profiles/rNN_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = re.compile(r"^(UTC\w+(\.\w+)*|LDTM\S*|STTM\S*|UBLKCP\S*|UTMA\S*|SYNCS\S*|ACQBULK|UCGABAR\S*|HMMA\S*|MUFU\S*|FFMA2|FADD2|FMUL2|"
                   r"LDG\.\S*256\S*|STG\.\S*256\S*|REDG\S*|ATOMG\S*|MATCH\S*)$")
COLLAPSE = (("SYNCS", "SYNCS"), ("LDTM", "LDTM"), ("STTM", "STTM"), ("UCGABAR", "UCGABAR"), ("UTCATOMSWS", "UTCATOMSWS"),
            ("LDG.", "LDG.256"), ("STG.", "STG.256"), ("REDG", "REDG"), ("ATOMG", "ATOMG"), ("MATCH", "MATCH"))


def main(paths):
    print("# cuobjdump -sass (sm_100a): instruction counts per kernel for the mnemonics that identify the Blackwell paths:")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA engine),")
    print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FFMA2/FADD2/FMUL2 = packed fp32, MUFU.TANH = tanh.approx.f32,")
    print("# LDG/STG.E.256 = 256-bit global accesses.  HMMA (legacy mma.sync) would be listed if there were any.")
    for path in paths:
        out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        kernels, cur = collections.OrderedDict(), None
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = kernels.setdefault(m.group(1), collections.Counter())
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m and cur is not None and WATCH.match(m.group(1)):
                name = m.group(1)
                for prefix, short in COLLAPSE:
                    if name.startswith(prefix):
                        name = short
                cur[name] += 1
        print("## %s" % os.path.relpath(path, ROOT))
        for k, c in sorted(kernels.items()):
            if c:
                dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
                dem = re.sub(r"\(.*", "", dem.replace("(anonymous namespace)::", "").replace("void ", "").replace("dsp::", ""))
                print("%-60s %s" % (dem[:60], ", ".join("%s=%d" % kv for kv in sorted(c.items()))))


if __name__ == "__main__":
    main(sys.argv[1:] or [os.path.join(ROOT, "deepsignal_plant_b200", "build", f) for f in ("kernels_tc.o", "freq.o", "comm.o", "extract.o")])
