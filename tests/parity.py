"""Shared helpers of the parity tests: case table, tolerance rule, GPU-vs-oracle comparison."""
import numpy as np

from meshlesshydro_b200 import capi, ic as IC
from cpu_oracles import Oracle, make_config as orc_config

# Floating-point tolerance of north_star: <= 1e-10 relative per particle.  "Relative" is taken
# against |reference value| + the field's scale (max |reference| over all particles): a pure
# per-value relative error is undefined where a gradient or a flux sum cancels to ~0.
RTOL = 1e-10


def close(a, b, rtol=RTOL, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "%s: NaN pattern differs (%d vs %d)" % (what, nan_a.sum(), nan_b.sum())
    fin = ~nan_a
    if not fin.any():
        return 0.0
    scale = np.max(np.abs(b[fin]))
    err = np.abs(a[fin] - b[fin]) / (np.abs(b[fin]) + scale + 1e-300)
    worst = float(err.max())
    assert worst <= rtol, "%s: max scaled error %.3e > %.1e" % (what, worst, rtol)
    return worst


# name -> (ic factory, preset, extra config)
CASES = {
    "kh_random_50": (lambda: IC.kelvin_helmholtz(50, lattice=False), "kh2d"),
    "kh_lattice_100": (lambda: IC.kelvin_helmholtz(100, lattice=True), "kh2d"),       # BASELINE config 1
    "kh_jitter_64": (lambda: IC.kelvin_helmholtz(64, lattice=True, jitter=0.2), "kh2d"),
    "fb_lattice_64": (lambda: IC.fluid_block(64), "fb2d"),
    "fb_jitter_60": (lambda: IC.fluid_block(60, jitter=0.05), "fb2d"),
    "sedov_21": (lambda: IC.sedov(21), "sedov3d"),
    "sedov_lattice_16": (lambda: IC.sedov(16, jitter=0.0), "sedov3d"),
}


def make_pair(case, abs_mode, max_ni=None, **gpu_over):
    """Build (ic, oracle, gpu) for a case with identical parameters."""
    factory, preset = CASES[case]
    ic = factory()
    ocfg = orc_config(preset, ic["h"], ic["gamma"], ic.get("box"), abs_mode=abs_mode)
    gcfg = capi.make_config(preset, ic["h"], ic["gamma"], ic.get("box"), abs_mode=abs_mode, debug_capture=1,
                            max_interactions=max_ni or 160, **gpu_over)
    orc = Oracle(ocfg, ic)
    gpu = capi.MfvGpu(gcfg)
    gpu.upload(ic)
    return ic, orc, gpu


def list_rows(flat, counts, cap):
    rows = flat.reshape(len(counts), cap)
    return [rows[i, :counts[i]] for i in range(len(counts))]


def compare_prepare(ic, orc, gpu):
    """After orc.step(stop_after=1) and gpu.prepare(): every intermediate up to the limited gradients."""
    D = ic["dim"]
    N = len(ic["x"])
    # --- integer / index work: bit exact ---
    assert np.array_equal(gpu.fetch("cell"), orc.fetch("cell")), "cell assignment differs"
    noi_g, noi_o = gpu.fetch("noi"), orc.fetch("noi")
    assert np.array_equal(noi_g, noi_o), "noi differs at %d particles" % int((noi_g != noi_o).sum())
    cap_g, cap_o = gpu.cfg.max_interactions, orc.cfg.max_ni
    rows_g = gpu.fetch("nnl").reshape(N, cap_g)
    rows_o = orc.fetch("nnl").reshape(N, cap_o)
    W = min(cap_g, cap_o)
    assert noi_o.max() <= W, "list capacity too small for the comparison"
    mask = np.arange(W)[None, :] < noi_o[:, None]
    assert np.array_equal(rows_g[:, :W][mask], rows_o[:, :W][mask]), "neighbour lists (order included) differ"
    if ic["periodic"]:
        ng_g, ng_o = gpu.fetch("noiGhosts"), orc.fetch("noiGhosts")
        assert np.array_equal(ng_g, ng_o), "noiGhosts differs at %d particles" % int((ng_g != ng_o).sum())
        parent = orc.fetch("ghost_parent")
        gl_o = orc.fetch("nnlGhosts").reshape(N, orc.cfg.max_gi)
        gl_g = gpu.fetch("nnlGhosts").reshape(N, cap_g)
        for i in np.nonzero(ng_o)[0]:
            assert np.array_equal(parent[gl_o[i, :ng_o[i]]], gl_g[i, :ng_o[i]]), "ghost list of particle %d differs" % i
    # --- floating point ---
    worst = {}
    for name in ("omega", "rho", "P"):
        worst[name] = close(gpu.fetch(name), orc.fetch(name), what=name)
    worst["Binv"] = close(gpu.fetch("Binv"), orc.fetch("Binv"), what="Binv")
    worst["gradPre"] = close(gpu.fetch("gradPre"), orc.fetch("gradPre"), what="gradPre")
    names = ["rhoGrad", "vxGrad", "vyGrad", "PGrad"] + (["vzGrad"] if D == 3 else [])
    for name in names:
        worst[name] = close(gpu.fetch(name), orc.fetch(name), what=name)
    return worst


def compare_state(ic, orc, gpu, rtol=RTOL, skip=None):
    D = ic["dim"]
    st = gpu.download_state()
    worst = {}
    keep = slice(None) if skip is None else ~skip
    for name in ["x", "y", "vx", "vy", "m", "u"] + (["z", "vz"] if D == 3 else []):
        worst[name] = close(st[name][keep], orc.fetch(name)[keep], rtol=rtol, what=name)
    return worst
