"""Shared helpers of the parity tests: case table, tolerance rule, GPU-vs-oracle comparison."""
import numpy as np

from meshlesshydro_b200 import capi, ic as IC
from cpu_oracles import Oracle, make_config as orc_config

# Floating-point tolerance of north_star: <= 1e-10 relative error PER PARTICLE.
#   * strictly positive per-particle quantities (x, y, z, m, u, rho, P, omega) are compared with the plain relative
#     error |gpu - ref| / |ref| of every single particle (`close_rel`).  The only concession: a value more than 14
#     decades below the field's maximum is measured against that floor instead of against itself.  Coordinates are
#     positions relative to an arbitrary origin: a particle that happens to sit within 1e-6 box lengths of a
#     coordinate plane is measured against 1e-6 of the box (its x = 3e-9 has no more significant digits than x = 0.3);
#   * quantities that are sums of cancelling terms (velocities of a fluid at rest, gradients, flux sums) have no
#     meaningful per-value relative error where they cancel to ~0; they are measured against |ref| + the field's
#     scale (max |ref| over all particles) (`close`).
RTOL = 1e-10
STRICT = ("x", "y", "z", "m", "u", "rho", "P", "omega")


def _finite_pair(a, b, what):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "%s: NaN pattern differs (%d vs %d)" % (what, nan_a.sum(), nan_b.sum())
    fin = ~nan_a
    return a[fin], b[fin]


def close(a, b, rtol=RTOL, what=""):
    """scaled rule: |a - b| / (|b| + max|b|) -- for cancelling quantities only"""
    a, b = _finite_pair(a, b, what)
    if a.size == 0:
        return 0.0
    scale = np.max(np.abs(b))
    err = np.abs(a - b) / (np.abs(b) + scale + 1e-300)
    worst = float(err.max())
    assert worst <= rtol, "%s: max scaled error %.3e > %.1e" % (what, worst, rtol)
    return worst


def close_rel(a, b, rtol=RTOL, what="", floor_frac=1e-14):
    """plain per-particle relative error |a - b| / |b| (floor: floor_frac of the field's maximum)"""
    a, b = _finite_pair(a, b, what)
    if a.size == 0:
        return 0.0
    floor = floor_frac * np.max(np.abs(b)) + 1e-300
    err = np.abs(a - b) / np.maximum(np.abs(b), floor)
    worst = float(err.max())
    assert worst <= rtol, "%s: max per-particle relative error %.3e > %.1e (particle %d: %r vs %r)" % (
        what, worst, rtol, int(err.argmax()), a[err.argmax()], b[err.argmax()])
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
    cap_g, cap_o = gpu.cfg.max_interactions, orc.max_ni
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
        gl_o = orc.fetch("nnlGhosts").reshape(N, orc.max_gi)
        gl_g = gpu.fetch("nnlGhosts").reshape(N, cap_g)
        for i in np.nonzero(ng_o)[0]:
            assert np.array_equal(parent[gl_o[i, :ng_o[i]]], gl_g[i, :ng_o[i]]), "ghost list of particle %d differs" % i
    # --- floating point ---
    worst = {}
    for name in ("omega", "rho", "P"):
        worst[name] = close_rel(gpu.fetch(name), orc.fetch(name), what=name)
    try:
        worst["Binv"] = close(gpu.fetch("Binv"), orc.fetch("Binv"), what="Binv")
    except KeyError:
        pass  # the reference-source build keeps only the psi-tilde weights, not the matrix
    worst["gradPre"] = close(gpu.fetch("gradPre"), orc.fetch("gradPre"), what="gradPre")
    names = ["rhoGrad", "vxGrad", "vyGrad", "PGrad"] + (["vzGrad"] if D == 3 else [])
    for name in names:
        worst[name] = close(gpu.fetch(name), orc.fetch(name), what=name)
    return worst


def compare_state(ic, orc, gpu, rtol=RTOL, skip=None, scaled=()):
    """`scaled`: fields of STRICT that fall back to the scaled rule for this case (say why at the call site)"""
    D = ic["dim"]
    st = gpu.download_state()
    worst = {}
    keep = slice(None) if skip is None else ~skip
    for name in ["x", "y", "vx", "vy", "m", "u"] + (["z", "vz"] if D == 3 else []):
        if name in ("x", "y", "z"):
            worst[name] = close_rel(st[name][keep], orc.fetch(name)[keep], rtol=rtol, what=name, floor_frac=1e-6)
        else:
            cmp = close_rel if (name in STRICT and name not in scaled) else close
            worst[name] = cmp(st[name][keep], orc.fetch(name)[keep], rtol=rtol, what=name)
    return worst


def face_reference(ic, orc, gpu, fields=("Aij", "WijR", "WijL", "vFrame")):
    """Call after orc.step(stop_after=1) and gpu.prepare(): maps every GPU face to the reference's slot of its canonical
    (lower-index) endpoint and stashes the reference's per-slot Aij (Particles.cpp:1290-1311) and WijR / WijL / vFrame
    (:1488-1733; ghosts :2504-2683) of those slots.  They must be read BEFORE the solve: Riemann::Riemann rotates the
    velocities of WijR / WijL in place (Riemann.cpp:19-81).  Faces of one-sided seam pairs (quirk Q9: the reference
    reads stale memory there) are skipped and counted."""
    D, N = ic["dim"], len(ic["x"])
    NW = D + 2
    pairs = gpu.fetch("face_pairs").reshape(-1, 3).astype(np.int64)
    nf = len(pairs)
    assert nf == int(gpu.fetch("num_faces")[0]) and nf > 0
    a, b, code = pairs[:, 0], pairs[:, 1], pairs[:, 2]
    assert np.all(a < b), "canonical endpoint must be the lower original index (quirk Q4)"
    M = orc.max_ni
    noi = orc.fetch("noi").astype(np.int64)
    nnl = orc.fetch("nnl").reshape(N, M).astype(np.int64)
    smask = np.arange(M)[None, :] < noi[:, None]
    rows = np.broadcast_to(np.arange(N, dtype=np.int64)[:, None], (N, M))[smask]
    keys = rows * N + nnl[smask]
    flat = (rows * M + np.broadcast_to(np.arange(M, dtype=np.int64)[None, :], (N, M))[smask])
    order = np.argsort(keys, kind="stable")
    keys, flat = keys[order], flat[order]
    reg = code == 0
    want = a[reg] * N + b[reg]
    pos = np.searchsorted(keys, want)
    assert np.all(pos < len(keys)) and np.array_equal(keys[np.minimum(pos, len(keys) - 1)], want), \
        "a regular face is missing from the reference's list of its canonical endpoint"
    sel = {"reg": (np.nonzero(reg)[0], flat[pos], "")}
    skipped = 0
    if ic["periodic"] and (~reg).any():
        MG = orc.max_gi
        parent = orc.fetch("ghost_parent").astype(np.int64)
        nog = orc.fetch("noiGhosts").astype(np.int64)
        gl = orc.fetch("nnlGhosts").reshape(N, MG).astype(np.int64)
        gmask = np.arange(MG)[None, :] < nog[:, None]
        grow = np.broadcast_to(np.arange(N, dtype=np.int64)[:, None], (N, MG))[gmask]
        gkeys = grow * N + parent[gl[gmask]]
        gflat = grow * MG + np.broadcast_to(np.arange(MG, dtype=np.int64)[None, :], (N, MG))[gmask]
        o2 = np.argsort(gkeys, kind="stable")
        gkeys, gflat = gkeys[o2], gflat[o2]
        gi = np.nonzero(~reg)[0]
        gwant = a[gi] * N + b[gi]
        gpos = np.minimum(np.searchsorted(gkeys, gwant), len(gkeys) - 1)
        found = gkeys[gpos] == gwant
        skipped = int((~found).sum())
        sel["ghost"] = (gi[found], gflat[gpos[found]], "Ghosts")
    stash = {}
    for kind, (fidx, slots, suffix) in sel.items():
        for name in fields:
            if name == "vFrame" and suffix:
                continue  # the ghost overload's frame velocities are not in the fetch list
            width = D if name in ("Aij", "vFrame") else NW
            stash[(kind, name)] = orc.fetch(name + suffix).reshape(-1, width)[slots].copy()
    # |grad f| of the limited gradients per particle, W order rho, P, vx, vy(, vz): scale of the reconstruction terms
    gnames = ["rhoGrad", "PGrad", "vxGrad", "vyGrad"] + (["vzGrad"] if D == 3 else [])
    gnorm = np.stack([np.sqrt((orc.fetch(n).reshape(N, D) ** 2).sum(axis=1)) for n in gnames], axis=1)
    return {"sel": sel, "stash": stash, "nf": nf, "skipped": skipped, "pairs": pairs, "gnorm": gnorm}


def compare_faces(ic, orc, gpu, pre, rtol=RTOL):
    """After the full step on both sides: the GPU's per-face record against the stash of face_reference and F against
    the reference's Fij (Particles.cpp:1787-1911) at the canonical slot.  Error measures (per face): A against |A|;
    the reconstructed + predicted states W = f + grad f . dx - dt/2 (..) (Particles.cpp:1609-1721) are sums that cancel
    where a cold particle sits next to the blast (|grad P| h >> P), so each component is measured against the larger of
    the two sides' values plus the size of the reconstruction terms (|grad f_a| + |grad f_b|) h/2 (velocities: plus the
    face's sound speed); F against the largest flux component of the face."""
    D = ic["dim"]
    NW = D + 2
    rec = gpu.fetch("face_rec").reshape(-1, 4 * D + 4)
    Fg = gpu.fetch("face_F").reshape(-1, NW)
    assert len(rec) == pre["nf"] == len(Fg)
    Wa, Wb, vF, A = rec[:, :NW], rec[:, NW:2 * NW], rec[:, 2 * NW:2 * NW + D], rec[:, 2 * NW + D:]
    worst = {}

    def upd(name, err):
        # entries where either side is non-finite (quirk Q13 can make the reference itself produce inf/NaN states) come
        # out as NaN: there both sides must show the same pattern, which the flux sums / state comparisons check
        if err.size:
            err = np.where(np.isnan(err), 0.0, err)
            worst[name] = max(worst.get(name, 0.0), float(err.max()))

    for kind, (fidx, slots, suffix) in pre["sel"].items():
        if fidx.size == 0:
            continue
        st = pre["stash"]
        refA = st[(kind, "Aij")]
        upd("Aij", np.abs(A[fidx] - refA).max(axis=1) / (np.sqrt((refA * refA).sum(axis=1)) + 1e-300))
        rR, rL = st[(kind, "WijR")], st[(kind, "WijL")]
        cs = np.sqrt(ic["gamma"] * np.maximum(np.abs(rR[:, 1] / rR[:, 0]), np.abs(rL[:, 1] / rL[:, 0])))
        ia, ib = pre["pairs"][fidx, 0], pre["pairs"][fidx, 1]
        recon = (pre["gnorm"][ia] + pre["gnorm"][ib]) * (0.5 * ic["h"])  # [face, component]
        for name, g, ref in (("WijR", Wa[fidx], rR), ("WijL", Wb[fidx], rL)):
            for c, cn in ((0, "rho"), (1, "P")):
                scale = np.maximum(np.abs(rR[:, c]), np.abs(rL[:, c])) + recon[:, c]
                upd(name + "." + cn, np.abs(g[:, c] - ref[:, c]) / scale)
            vscale = np.maximum(np.abs(rR[:, 2:]).max(axis=1), np.abs(rL[:, 2:]).max(axis=1)) + cs + recon[:, 2:].max(axis=1)
            upd(name + ".v", np.abs(g[:, 2:] - ref[:, 2:]).max(axis=1) / vscale)
        if (kind, "vFrame") in st:
            ref = st[(kind, "vFrame")]
            upd("vFrame", np.abs(vF[fidx] - ref).max(axis=1) / (np.abs(ref).max(axis=1) + cs))
        ref = orc.fetch("Fij" + suffix).reshape(-1, NW)[slots]
        upd("Fij", np.abs(Fg[fidx] - ref).max(axis=1) / (np.abs(ref).max(axis=1) + 1e-300))
        del ref
    bad = {k: v for k, v in worst.items() if not (v <= rtol)}
    assert not bad, "per-face intermediates off: %r (tolerance %.1e)" % (bad, rtol)
    worst["faces"] = pre["nf"]
    worst["skipped_one_sided"] = pre["skipped"]
    return worst
