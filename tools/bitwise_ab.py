"""Hashes of the state and of the per-particle intermediates after a few steps, to check that a kernel change that is
meant to be bit-neutral is: run once per library build (MLH_GPU_LIB=...) and compare the lines."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshlesshydro_b200 import capi, ic as IC

cases = [("kh_jitter_300", IC.kelvin_helmholtz(300, lattice=True, jitter=0.2), "kh2d", {}),
         ("kh_random_100", IC.kelvin_helmholtz(100), "kh2d", {}),
         ("fb_jitter_200", IC.fluid_block(200, jitter=0.05), "fb2d", {}),
         ("sedov_41", IC.sedov(41), "sedov3d", dict(q13_mode=capi.Q13_ZERO_Z, max_interactions=128))]
for name, ic, preset, over in cases:
    cfg = capi.make_config(preset, ic["h"], ic["gamma"], ic.get("box"), abs_mode=capi.ABS_INT_TRUNC, debug_capture=1, **over)
    g = capi.MfvGpu(cfg); g.upload(ic)
    for _ in range(3): g.step()
    h = hashlib.sha256()
    st = g.download_state()
    for k in sorted(st):
        if st[k] is not None: h.update(np.ascontiguousarray(st[k]).tobytes())
    for f in ("omega", "rho", "Binv", "rhoGrad", "PGrad", "mF", "eF", "vF"):
        h.update(np.ascontiguousarray(g.fetch(f)).tobytes())
    print(name, len(ic["x"]), "flags", g.error_flags(), h.hexdigest()[:24])
    g.close()
