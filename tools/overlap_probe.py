"""Probe: does the flux stage gain from running next to another stream's kernels?  Two contexts (own streams) step two
half-size KH problems concurrently; their aggregate rate is compared with one context on the full problem.  Grid sizes of
the face kernels come from MLH_GRID_* (a smaller resident footprint lets kernels of the two streams share an SM).
usage: python tools/overlap_probe.py [side_full]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshlesshydro_b200 import capi, ic as IC

def ctx(side):
    ic = IC.kelvin_helmholtz(side, lattice=True, jitter=0.2)
    cfg = capi.make_config("kh2d", ic["h"], ic["gamma"], ic.get("box"), abs_mode=capi.ABS_INT_TRUNC, max_interactions=96)
    g = capi.MfvGpu(cfg); g.upload(ic)
    return g, len(ic["x"])

def timed(gs, steps):
    for g in gs:
        for _ in range(2): g.step(want_dt=False)
    for g in gs: g.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for g in gs: g.step(want_dt=False)
    for g in gs: g.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3

side = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
half = int(round(side / 2 ** 0.5))
g, n = ctx(side)
ms = timed([g], 5)
print("single  N=%d  %.3f ms/step  %.4e particle-updates/s" % (n, ms, n / ms * 1e3)); g.close()
g1, n1 = ctx(half); g2, n2 = ctx(half)
ms1 = timed([g1], 5)
print("half    N=%d  %.3f ms/step  %.4e particle-updates/s" % (n1, ms1, n1 / ms1 * 1e3))
ms2 = timed([g1, g2], 5)
print("dual    N=%d+%d  %.3f ms per pair of steps  %.4e particle-updates/s  (serial would be %.3f)" % (n1, n2, ms2, (n1 + n2) / ms2 * 1e3, 2 * ms1))
