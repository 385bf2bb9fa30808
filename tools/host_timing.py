import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np
from meshlesshydro_b200 import capi, ic as IC
which = sys.argv[1]
if len(sys.argv) > 2 and sys.argv[2] == "torch":
    import torch; torch.cuda.set_device(0); torch.zeros(1, device="cuda")
ic = IC.sedov(61) if which == "sedov61" else IC.kelvin_helmholtz(1000, lattice=True, jitter=0.2)
preset = "sedov3d" if which == "sedov61" else "kh2d"
cfg = capi.make_config(preset, ic["h"], ic["gamma"], ic.get("box"), abs_mode=capi.ABS_INT_TRUNC, max_interactions=128 if ic["dim"] == 3 else 96)
g = capi.MfvGpu(cfg); g.upload(ic)
for _ in range(3): g.step(want_dt=False)
g.synchronize()
t0 = time.perf_counter()
for _ in range(20): g.step(want_dt=False)
t1 = time.perf_counter(); g.synchronize(); t2 = time.perf_counter()
print(which, sys.argv[2:], "host enqueue of 20 steps %.2f ms/step, incl. GPU %.2f ms/step" % ((t1 - t0) * 50, (t2 - t0) * 50))
