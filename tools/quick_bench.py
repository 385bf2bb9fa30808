"""Scratch timing of the step on one GPU with the per-kernel CUDA-event profile (not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshlesshydro_b200 import capi, ic as IC

def run(name, ic, preset, steps=5, **over):
    over.setdefault("abs_mode", capi.ABS_INT_TRUNC)
    cfg = capi.make_config(preset, ic["h"], ic["gamma"], ic.get("box"), **over)
    g = capi.MfvGpu(cfg); g.upload(ic)
    N = len(ic["x"])
    for _ in range(2): g.step(want_dt=False)
    g.synchronize()
    g.timer_start()
    for _ in range(steps): g.step(want_dt=False)
    ms = g.timer_stop() / steps
    g.profile(True)
    for _ in range(steps): g.step(want_dt=False)
    prof = g.profile_read(); g.profile(False)
    noi = g.fetch("noi"); ng = g.fetch("noiGhosts")
    print(f"{name}: N={N} {ms:.3f} ms/step -> {N/ms*1e3:.3e} particle-updates/s  K={noi.mean():.1f}+{ng.mean():.2f} flags={g.error_flags()} sums={g.sums()}")
    tot = sum(v[0] for v in prof.values())
    for k,(t,l) in prof.items():
        if l: print(f"    {k:24s} {t/steps:9.3f} ms/step  {100*t/tot:5.1f}%  launches/step={l/steps:.0f}")
    g.close()

which = sys.argv[1:] or ["kh100","sedov61","fb1000"]
for w in which:
    if w == "kh100": run(w, IC.kelvin_helmholtz(100), "kh2d")
    if w == "kh500": run(w, IC.kelvin_helmholtz(500), "kh2d")
    if w == "kh1000": run(w, IC.kelvin_helmholtz(1000), "kh2d")
    if w == "kh1000j": run(w, IC.kelvin_helmholtz(1000, lattice=True, jitter=0.2), "kh2d", max_interactions=96)
    if w == "kh2000": run(w, IC.kelvin_helmholtz(2000), "kh2d", steps=3)
    if w == "kh2000j": run(w, IC.kelvin_helmholtz(2000, lattice=True, jitter=0.2), "kh2d", steps=3, max_interactions=96)
    if w == "fb1000j": run(w, IC.fluid_block(1000, jitter=0.05), "fb2d", max_interactions=96)
    if w == "sedov31": run(w, IC.sedov(31), "sedov3d")
    if w == "sedov61": run(w, IC.sedov(61), "sedov3d", abs_mode=capi.ABS_INT_TRUNC, q13_mode=capi.Q13_ZERO_Z, max_interactions=128)
    if w == "sedov128": run(w, IC.sedov(128), "sedov3d", steps=3)
    if w == "fb1000": run(w, IC.fluid_block(1000), "fb2d")
