"""Generate tests/golden/*.npz from the REFERENCE-SOURCE oracle (oracle/_ref, i.e. the reference's own
Particles/Domain/Riemann/Helper sources compiled by oracle/ref_build/Makefile).

Run in the build container (needs /root/reference to have built oracle/_ref):
    python tools/make_golden.py
Each fixture holds the seeded input particles, the run parameters and the reference's outputs after
`prepare` (cells, ordered neighbour lists, omega, rho, P, pre-/post-limiter gradients, dt_cfl) and after
one full step (per-particle flux sums and the updated state), plus the conservation sums.  The GPU box
has no /root/reference; tests there compare against these files (tests/test_golden*.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from meshlesshydro_b200 import ic as IC  # noqa: E402
from cpu_oracles import Reference  # noqa: E402

# name -> (ic factory, oracle/_ref variant, preset, abs_mode)
FIXTURES = {
    "kh_random_24_inttrunc": (lambda: IC.kelvin_helmholtz(24, lattice=False), "kh2d", "kh2d", 0),
    "kh_random_24_fabs": (lambda: IC.kelvin_helmholtz(24, lattice=False), "kh2d_fabs", "kh2d", 1),
    "kh_lattice_32_fabs": (lambda: IC.kelvin_helmholtz(32, lattice=True), "kh2d_fabs", "kh2d", 1),
    "fb_jitter_28_inttrunc": (lambda: IC.fluid_block(28, jitter=0.05), "fb2d", "fb2d", 0),
    "fb_lattice_24_fabs": (lambda: IC.fluid_block(24), "fb2d_fabs", "fb2d", 1),
    "sedov_10_inttrunc": (lambda: IC.sedov(10), "sedov3d", "sedov3d", 0),
    "sedov_10_fabs": (lambda: IC.sedov(10), "sedov3d_fabs", "sedov3d", 1),
}


def compact_lists(flat, counts, cap):
    rows = flat.reshape(len(counts), cap)
    return np.concatenate([rows[i, :counts[i]] for i in range(len(counts))]).astype(np.int32)


def make(name):
    factory, variant, preset, abs_mode = FIXTURES[name]
    ic = factory()
    D = ic["dim"]
    out = dict(preset=preset, variant=variant, abs_mode=abs_mode, dim=D, periodic=ic["periodic"], h=ic["h"],
               gamma=ic["gamma"], box=np.zeros(0) if ic.get("box") is None else ic["box"])
    for k in ("x", "y", "z", "vx", "vy", "vz", "m", "u"):
        if ic.get(k) is not None:
            out["in_" + k] = ic[k]
    ref = Reference(variant, ic)
    out["sums0"] = ref.sums()
    dt = ref.step(stop_after=1)
    out["dt_cfl"] = dt
    cells, cs, bounds = ref.grid()
    out["cells"], out["cell_size"], out["bounds"] = cells, cs, bounds
    noi = ref.fetch("noi")
    out["cell"], out["noi"] = ref.fetch("cell"), noi
    out["nnl"] = compact_lists(ref.fetch("nnl"), noi, ref.info["max_ni"])
    if ic["periodic"]:
        ng = ref.fetch("noiGhosts")
        parent = ref.fetch("ghost_parent")
        out["noiGhosts"] = ng
        out["nnlGhostParents"] = parent[compact_lists(ref.fetch("nnlGhosts"), ng, ref.info["max_gi"])].astype(np.int32)
    for k in ["omega", "rho", "P", "gradPre", "rhoGrad", "vxGrad", "vyGrad", "PGrad"] + (["vzGrad"] if D == 3 else []):
        out[k] = ref.fetch(k)
    ref2 = Reference(variant, ic)
    ref2.step(dt_fixed=dt)
    for k in ["mF", "eF", "vF", "x", "y", "vx", "vy", "m", "u"] + (["z", "vz"] if D == 3 else []):
        out["out_" + k] = ref2.fetch(k)
    out["sums1"] = ref2.sums()
    return out


if __name__ == "__main__":
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    for name in (sys.argv[1:] or FIXTURES):
        data = make(name)
        path = os.path.join(gdir, name + ".npz")
        np.savez_compressed(path, **data)
        print("%s: N=%d dt=%.6e  %d kB" % (name, len(data["in_x"]), data["dt_cfl"], os.path.getsize(path) // 1024))
