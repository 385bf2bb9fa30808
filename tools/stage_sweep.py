"""Sweep of mlh_config.stage_bytes (faces per K4 chunk): does an L2-resident staging buffer pay for the extra launches?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshlesshydro_b200 import capi, ic as IC

def run(name, ic, preset, stage, steps=5, **over):
    cfg = capi.make_config(preset, ic["h"], ic["gamma"], ic.get("box"), abs_mode=capi.ABS_INT_TRUNC, **over)
    cfg.stage_bytes = stage
    g = capi.MfvGpu(cfg); g.upload(ic)
    N = len(ic["x"])
    for _ in range(2): g.step(want_dt=False)
    g.synchronize(); g.timer_start()
    for _ in range(steps): g.step(want_dt=False)
    ms = g.timer_stop() / steps
    g.profile(True)
    for _ in range(steps): g.step(want_dt=False)
    prof = g.profile_read(); g.profile(False)
    k4 = {k: (round(t / steps, 3), int(l / steps)) for k, (t, l) in prof.items() if k.startswith("k4")}
    print(f"{name} stage={stage>>20} MiB: {ms:.3f} ms/step  {k4}")
    g.close()

sed = IC.sedov(61); kh = IC.kelvin_helmholtz(1000, lattice=True, jitter=0.2)
for mb in (0, 32, 64, 96, 128, 256, 512):
    run("sedov61", sed, "sedov3d", mb << 20, q13_mode=capi.Q13_ZERO_Z, max_interactions=128)
for mb in (0, 64, 128, 256, 1024):
    run("kh1000j", kh, "kh2d", mb << 20, max_interactions=96)
