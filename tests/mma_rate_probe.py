"""Diagnostic (run on the GPU box): SM cycles per tcgen05.mma for several shapes / operand
sources / accumulator-chain interleavings, all SMs busy.  Not a test."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsignal_plant_b200 import _native
L = _native.lib()
for n_bits, n in ((64, 64), (0, 128), (2, 256)):
    for ts in (0, 1):
        for chains_log in (0, 1, 2):
            if n * (1 << chains_log) > 384 - (0 if not ts else 0):
                continue
            r = C.c_double()
            _native.check(L.dsp_selftest(0, 100 + n_bits + ts + (chains_log << 4), C.byref(r)))
            print("N=%-3d %s chains=%d : %6.1f cycles/MMA  -> %5.1f%% of 4096 MAC/clk" % (
                n, "TS" if ts else "SS", 1 << chains_log, r.value, 100 * (128 * n * 16 / r.value) / 4096))
