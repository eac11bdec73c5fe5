"""Diagnostic (run on the GPU box): do MUFU and packed / scalar FMA instructions overlap on an SM?
16 warps per SM, 8 independent chains per thread; prints SM cycles per loop iteration.  Not a test."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsignal_plant_b200 import _native
L = _native.lib()
names = {1: "8 MUFU.EX2", 8: "8 MUFU.RCP", 2: "8 FFMA2", 4: "16 FFMA (scalar)", 3: "8 EX2 + 8 FFMA2", 5: "8 EX2 + 16 FFMA",
         9: "8 EX2 + 8 RCP", 11: "8 EX2 + 8 RCP + 8 FFMA2", 6: "8 FFMA2 + 16 FFMA"}
for mode in (1, 8, 9, 2, 4, 6, 3, 5, 11):
    r = C.c_double()
    _native.check(L.dsp_selftest(0, 300 + mode, C.byref(r)))
    print("%-28s %7.1f cycles / iteration (16 warps = 4 per scheduler)" % (names[mode], r.value))
