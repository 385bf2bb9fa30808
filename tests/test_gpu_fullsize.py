"""GPU (-m gpu): the BASELINE.json configurations at FULL size, checked through size-independent properties (the CPU
oracle needs minutes to hours there: ~6e4 particle-updates/s).

Properties (each follows from the reference's algorithm, not from this implementation):
  * cell sort: the sorted order is a permutation, cell ids are non-decreasing along it and inside a cell the
    ORIGINAL indices ascend (= push_back order of Cell::prtcls, Particles.cpp:319);
  * every cell id equals the reference's formula evaluated in numpy on the same coordinates (Particles.cpp:279-302);
  * neighbour relation symmetric (|x_i - x_j|^2 is computed from bit-identical operands on both sides), no self
    entries, list length == noi, and #faces == #pairs / 2 (each pair evaluated once, quirk Q4);
  * every listed pair is inside the kernel support and -- on a sample of particles -- the list equals the brute-force
    neighbour set of the reference's cutoff test `pow(dx,2)+pow(dy,2)[+pow(dz,2)] < h*h` (bit-exact, order included);
  * total mass / momentum / energy conserved to summation round-off over several adaptive steps, no device flags.
"""
import numpy as np
import pytest

from meshlesshydro_b200 import capi, ic as IC

pytestmark = pytest.mark.gpu

# name -> (factory, preset, max_interactions, steps, full neighbour-symmetry check)
FULL = {
    "sedov_61": (lambda: IC.sedov(61), "sedov3d", 128, 3, True),                                           # configs[1]
    "fb_1000": (lambda: IC.fluid_block(1000, jitter=0.05), "fb2d", 96, 3, True),                           # configs[2]
    "kh_2000": (lambda: IC.kelvin_helmholtz(2000, lattice=True, jitter=0.2), "kh2d", 96, 2, False),        # configs[3]
}


def _reference_cells(ic, gpu):
    cells, size, bounds = gpu.grid()
    D = ic["dim"]
    cid = np.zeros(len(ic["x"]), dtype=np.int64)
    stride = 1
    for k, name in enumerate(("x", "y", "z")[:D]):
        f = np.floor((ic[name] - bounds[k]) / size[k]).astype(np.int64)
        f[f == cells[k]] -= 1
        cid += f * stride
        stride *= int(cells[k])
    return cid


@pytest.mark.parametrize("name", list(FULL))
def test_full_size_properties(name):
    factory, preset, max_ni, steps, full_sym = FULL[name]
    ic = factory()
    D, N = ic["dim"], len(ic["x"])
    cfg = capi.make_config(preset, ic["h"], ic["gamma"], ic.get("box"), abs_mode=capi.ABS_INT_TRUNC, max_interactions=max_ni)
    gpu = capi.MfvGpu(cfg)
    gpu.upload(ic)
    dt0 = gpu.prepare()
    assert gpu.error_flags() == 0 and dt0 > 0.0

    # ---- cell sort ----
    cell = gpu.fetch("cell").astype(np.int64)
    assert np.array_equal(cell, _reference_cells(ic, gpu)), "cell ids differ from the reference formula"
    pos = gpu.fetch("sorted_index").astype(np.int64)  # original index -> position in the sorted set
    order = np.empty(N, dtype=np.int64)
    order[pos] = np.arange(N)
    assert np.array_equal(np.sort(pos), np.arange(N)), "sorted order is not a permutation"
    cs = cell[order]
    assert np.all(np.diff(cs) >= 0), "cells are not sorted"
    same = np.diff(cs) == 0
    assert np.all(np.diff(order)[same] > 0), "original indices do not ascend inside a cell"

    # ---- neighbour lists ----
    noi = gpu.fetch("noi").astype(np.int64)
    nog = gpu.fetch("noiGhosts").astype(np.int64)
    nfaces = int(gpu.fetch("num_faces")[0])
    assert (noi.sum() + nog.sum()) % 2 == 0 and nfaces == (noi.sum() + nog.sum()) // 2, (noi.sum(), nog.sum(), nfaces)
    coords = np.stack([ic[k] for k in ("x", "y", "z")[:D]], axis=1)
    hsqr = ic["h"] * ic["h"]
    if full_sym:
        rows = gpu.fetch("nnl").reshape(N, max_ni)
        mask = np.arange(max_ni)[None, :] < noi[:, None]
        ii = np.broadcast_to(np.arange(N)[:, None], rows.shape)[mask].astype(np.int64)
        jj = rows[mask].astype(np.int64)
        assert not np.any(ii == jj), "self entry in a neighbour list"
        key = np.sort(ii * N + jj)
        assert np.all(np.diff(key) > 0), "duplicate neighbour entry"
        assert np.array_equal(key, np.sort(jj * N + ii)), "neighbour relation is not symmetric"
        d = coords[ii] - coords[jj]
        dsq = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
        if D == 3:
            dsq = dsq + d[:, 2] * d[:, 2]
        assert np.all(dsq < hsqr), "listed pair outside the kernel support"
        # brute force on a sample of particles: exact list (order = stencil cell order, ascending index inside a cell)
        rng = np.random.default_rng(7)
        cells, _, _ = gpu.grid()
        for i in rng.choice(N, size=24, replace=False):
            d = coords[i][None, :] - coords
            dsq = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
            if D == 3:
                dsq = dsq + d[:, 2] * d[:, 2]
            cand = np.nonzero(dsq < hsqr)[0]
            cand = cand[cand != i]
            assert set(cand.tolist()) == set(rows[i, :noi[i]].tolist()), "neighbour set of particle %d differs" % i
            # Domain::getNeighborCells order (Domain.cpp:83-118): x offset outer, then y, z inner
            ci = [int(cell[i] % cells[0]), int(cell[i] // cells[0] % cells[1]), int(cell[i] // (cells[0] * cells[1]))]
            cj = np.stack([cell[cand] % cells[0], cell[cand] // cells[0] % cells[1], cell[cand] // (cells[0] * cells[1])], axis=1)
            off = cj - np.array(ci)[None, :]
            rank = (off[:, 0] + 1) * 9 + (off[:, 1] + 1) * 3 + (off[:, 2] + 1 if D == 3 else 0)
            expect = cand[np.lexsort((cand, rank))]
            assert np.array_equal(expect, rows[i, :noi[i]]), "neighbour ORDER of particle %d differs" % i
        del rows, mask, ii, jj, key

    # ---- conservation over adaptive steps ----
    s0 = gpu.sums()
    gpu.advance(dt0)
    for _ in range(steps - 1):
        gpu.step()
    s1 = gpu.sums()
    assert gpu.error_flags() & ~capi.F_NEG_GHOST_PRESSURE == 0
    assert gpu.fetch("counters")[0] == 0, "one-sided seam pairs present"
    p_scale = np.sqrt(2.0 * s0[1] * s0[2])
    assert abs(s1[1] - s0[1]) <= 1e-12 * s0[1], ("mass", s0[1], s1[1])
    assert abs(s1[2] - s0[2]) <= 1e-11 * s0[2], ("energy", s0[2], s1[2])
    for k in range(3, 3 + D):
        assert abs(s1[k] - s0[k]) <= 1e-11 * p_scale, ("momentum", k, s0[k], s1[k])
    st = gpu.download_state()
    for k in ("x", "y", "vx", "vy", "m", "u"):
        assert np.all(np.isfinite(st[k])), k
    assert np.all(st["m"] > 0.0)
    print(name, "N=%d faces=%d mean K=%.2f dM/M=%.1e dE/E=%.1e" % (N, nfaces, (noi.mean() + nog.mean()),
                                                                    abs(s1[1] - s0[1]) / s0[1], abs(s1[2] - s0[2]) / s0[2]))
    gpu.close()
