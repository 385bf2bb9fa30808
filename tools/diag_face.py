"""Scratch diagnosis (GPU box): worst particle of a case in `u` after one step, and the per-face differences around it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from meshlesshydro_b200 import capi

case = sys.argv[1] if len(sys.argv) > 1 else "sedov_lattice_16"
abs_mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ic, orc, gpu = parity.make_pair(case, abs_mode)
D = ic["dim"]; NW = D + 2
dt = orc.step(stop_after=1); gpu.prepare()
pre = parity.face_reference(ic, orc, gpu)
orc.step(dt_fixed=dt); gpu.advance(dt)
st = gpu.download_state()
for name in ("u", "m"):
    ref = orc.fetch(name)
    err = np.abs(st[name] - ref) / np.abs(ref)
    w = int(err.argmax())
    print(name, "worst particle", w, "rel err %.3e" % err[w], st[name][w], ref[w])
ref = orc.fetch("u"); err = np.abs(st["u"] - ref) / np.abs(ref); w = int(err.argmax())
for nm in ("mF", "eF"):
    g, o = gpu.fetch(nm), orc.fetch(nm)
    print(nm, "gpu %.17e ref %.17e" % (g[w], o[w]))
pairs = pre["pairs"]
rec = gpu.fetch("face_rec").reshape(-1, 4 * D + 4); Fg = gpu.fetch("face_F").reshape(-1, NW)
fidx, slots, _ = pre["sel"]["reg"]
Fo = orc.fetch("Fij").reshape(-1, NW)[slots]
mine = np.nonzero((pairs[fidx, 0] == w) | (pairs[fidx, 1] == w))[0]
print("faces of particle", w, ":", len(mine))
for q in mine:
    f = fidx[q]
    dF = Fg[f] - Fo[q]
    rR, rL = pre["stash"][("reg", "WijR")][q], pre["stash"][("reg", "WijL")][q]
    print(" pair", pairs[f, :2], "F_E gpu %.6e ref %.6e diff %.2e | dW_R %.1e dW_L %.1e | PR %.10e PL %.10e" % (
        Fg[f, 1], Fo[q, 1], dF[1], np.abs(rec[f, :NW] - rR).max(), np.abs(rec[f, NW:2*NW] - rL).max(), rR[1], rL[1]))
