"""Loader for tests/golden/*.npz (made by tools/make_golden.py from oracle/_ref = the reference's own sources)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("preset", "variant"):
        g[k] = str(g[k])
    for k in ("abs_mode", "dim", "periodic"):
        g[k] = int(g[k])
    for k in ("h", "gamma", "dt_cfl"):
        g[k] = float(g[k])
    ic = dict(dim=g["dim"], periodic=g["periodic"], h=g["h"], gamma=g["gamma"],
              box=g["box"] if g["box"].size else None)
    for k in ("x", "y", "z", "vx", "vy", "vz", "m", "u"):
        ic[k] = g.get("in_" + k)
    return g, ic


def split_lists(flat, counts):
    return np.split(flat, np.cumsum(counts)[:-1])


def compact(flat, counts, cap):
    rows = flat.reshape(len(counts), cap)
    mask = np.arange(cap)[None, :] < counts[:, None]
    return rows[mask]
