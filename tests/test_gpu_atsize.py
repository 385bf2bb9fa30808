"""GPU (-m gpu): floating-point parity AT THE SIZES THE BENCHMARKS RUN.

* BASELINE configs[1], Sedov 3D N = 61^3: the CUDA path against the reference's OWN sources
  (oracle/_ref/libref_sedov3d_bench.so = Particles.cpp / Domain.cpp / Riemann.cpp / Helper.cpp compiled by
  oracle/ref_build/Makefile with MAX_NUM_INTERACTIONS 128), one full step: ~4 s of CPU per pass.
* BASELINE configs[2], fluid block 2D N = 1000^2: against oracle/liboracle.so (the C restatement, itself pinned
  bit for bit to the reference sources by tests/test_oracle_vs_ref.py) with 64 slots per particle: ~25 s per pass.
* Kelvin-Helmholtz 2D periodic N = 1000^2 against the oracle (~1 min); BASELINE configs[3], N = 2000^2 = the default
  bench workload, behind MLH_ATSIZE_KH2000=1 (~6 min of CPU).

Bars: cell ids and ordered neighbour lists bit-exact; omega, rho, P, x, m, u per-particle relative <= 1e-10;
gradients, flux sums, velocities <= 1e-10 under the scaled rule (tests/parity.py); per-face A_ij, W_L/W_R, F_ij
(Particles.cpp:1290-1911) <= 1e-10 at Sedov 61^3.
"""
import os

import numpy as np
import pytest

from meshlesshydro_b200 import capi, ic as IC
import parity
from cpu_oracles import Oracle, Reference, make_config as orc_config

pytestmark = pytest.mark.gpu


def _one_step(ic, orc, gpu, faces):
    dt_o = orc.step(stop_after=1)
    dt_g = gpu.prepare()
    assert gpu.error_flags() == 0
    assert abs(dt_g - dt_o) <= 1e-12 * dt_o, (dt_g, dt_o)
    cg, sg, bg = gpu.grid()
    co, so, bo = orc.grid()
    assert np.array_equal(cg, co) and np.array_equal(sg, so) and np.array_equal(bg, bo), "search grid differs"
    worst = parity.compare_prepare(ic, orc, gpu)
    pre = parity.face_reference(ic, orc, gpu) if faces else None
    orc.step(dt_fixed=dt_o)  # a stopped step leaves the state untouched: this is the same step, completed
    gpu.advance(dt_o)
    assert gpu.error_flags() == 0
    for name in ("mF", "eF", "vF"):
        worst[name] = parity.close(gpu.fetch(name), orc.fetch(name), what=name)
    worst.update(parity.compare_state(ic, orc, gpu))
    if faces:
        worst.update(parity.compare_faces(ic, orc, gpu, pre))
    return worst


def test_sedov_61_cubed_vs_reference_sources():
    variant = "sedov3d_bench"
    if not Reference.available(variant):
        pytest.skip("oracle/_ref/libref_%s.so not built (needs /root/reference at build time)" % variant)
    ic = IC.sedov(61)
    ref = Reference(variant, ic)
    assert ref.info["max_ni"] == 128 and not ref.info["fabs"]
    cfg = capi.make_config("sedov3d", ic["h"], ic["gamma"], None, abs_mode=capi.ABS_INT_TRUNC, q13_mode=capi.Q13_ZERO_Z,
                           debug_capture=1, max_interactions=128)
    gpu = capi.MfvGpu(cfg)
    gpu.upload(ic)
    worst = _one_step(ic, ref, gpu, faces=True)
    print("sedov 61^3 vs reference sources:", {k: ("%.1e" % v if isinstance(v, float) else v) for k, v in worst.items()})
    gpu.close()
    ref.close()


def test_fluid_block_1000_squared_vs_oracle():
    ic = IC.fluid_block(1000, jitter=0.05)
    ocfg = orc_config("fb2d", ic["h"], ic["gamma"], None, abs_mode=0, max_ni=64, max_gi=64)
    orc = Oracle(ocfg, ic)
    cfg = capi.make_config("fb2d", ic["h"], ic["gamma"], None, abs_mode=capi.ABS_INT_TRUNC, debug_capture=1,
                           max_interactions=64)
    gpu = capi.MfvGpu(cfg)
    gpu.upload(ic)
    worst = _one_step(ic, orc, gpu, faces=False)
    print("fluid block 1000^2 vs oracle:", {k: ("%.1e" % v if isinstance(v, float) else v) for k, v in worst.items()})
    gpu.close()
    orc.close()


def _kh_at_size(side):
    ic = IC.kelvin_helmholtz(side, lattice=True, jitter=0.2)  # the generator bench.py uses for kh1000 / kh2000
    ocfg = orc_config("kh2d", ic["h"], ic["gamma"], ic.get("box"), abs_mode=0, max_ni=96, max_gi=96)
    orc = Oracle(ocfg, ic)
    cfg = capi.make_config("kh2d", ic["h"], ic["gamma"], ic.get("box"), abs_mode=capi.ABS_INT_TRUNC, debug_capture=1,
                           max_interactions=96)
    gpu = capi.MfvGpu(cfg)
    gpu.upload(ic)
    worst = _one_step(ic, orc, gpu, faces=False)
    print("KH %d^2 (periodic) vs oracle:" % side, {k: ("%.1e" % v if isinstance(v, float) else v) for k, v in worst.items()})
    gpu.close()
    orc.close()


def test_kelvin_helmholtz_1000_squared_vs_oracle():
    """periodic 2D at 1 M particles: ghost lists, images and the seam at size (~1 min of CPU for the oracle)"""
    _kh_at_size(1000)


@pytest.mark.skipif(not os.environ.get("MLH_ATSIZE_KH2000"), reason="BASELINE configs[3] = bench.py's default workload, "
                    "one step against the oracle: ~6 min of CPU; set MLH_ATSIZE_KH2000=1 (log of the last run: profiles/)")
def test_kelvin_helmholtz_2000_squared_vs_oracle():
    _kh_at_size(2000)


# ---- the chunked flux pass (launch_chunks, csrc/k4_flux.cu): C4 on one GPU and C5 on eight run 2-8 staging chunks per
# step; the reference makes one pass over all slots (Particles.cpp:1787-1911), so chunking must not change a single bit
@pytest.mark.parametrize("case,stage_bytes", [("kh_jitter_64", 5 << 20), ("sedov_21", 8 << 20), ("fb_jitter_60", 3 << 20)])
def test_chunked_flux_pass_is_bitwise_identical(case, stage_bytes):
    states = []
    for sb in (0, stage_bytes):
        ic, orc, gpu = parity.make_pair(case, capi.ABS_INT_TRUNC, stage_bytes=sb)
        dts = [gpu.step() for _ in range(2)]
        nfaces = int(gpu.fetch("num_faces")[0])
        states.append((dts, gpu.download_state(), gpu.error_flags()))
        gpu.close()
    D = ic["dim"]
    per_face = (4 * D + 4 + 1 + 12) * 8 + 4  # record + P* + solver queue entry (mlh_stage_alloc)
    assert nfaces * per_face >= 3 * stage_bytes, "the small budget must force >= 3 chunks (%d faces)" % nfaces
    assert states[0][2] == states[1][2], "device flags differ between one chunk and many"
    assert states[0][0] == states[1][0], "dt differs between one chunk and many"
    for k, v in states[0][1].items():
        if v is not None:
            assert np.array_equal(v, states[1][1][k]), "%s differs between one chunk and many" % k


def test_chunked_flux_pass_matches_oracle():
    ic, orc, gpu = parity.make_pair("sedov_21", capi.ABS_FABS, stage_bytes=9 << 20)
    dt_o = orc.step()
    dt_g = gpu.step()
    assert abs(dt_g - dt_o) <= 1e-12 * dt_o
    for name in ("mF", "eF", "vF"):
        parity.close(gpu.fetch(name), orc.fetch(name), what=name)
    parity.compare_state(ic, orc, gpu)


# ---- per-face intermediates on the small cases, both abs modes, periodic faces included ----
@pytest.mark.parametrize("abs_mode", [capi.ABS_INT_TRUNC, capi.ABS_FABS])
@pytest.mark.parametrize("case", ["kh_random_50", "kh_jitter_64", "fb_jitter_60", "sedov_21"])
def test_per_face_intermediates_match_oracle(case, abs_mode):
    ic, orc, gpu = parity.make_pair(case, abs_mode)
    dt = orc.step(stop_after=1)
    gpu.prepare()
    pre = parity.face_reference(ic, orc, gpu)
    orc.step(dt_fixed=dt)
    gpu.advance(dt)
    w = parity.compare_faces(ic, orc, gpu, pre)
    if ic["periodic"]:
        assert w["skipped_one_sided"] == int(gpu.fetch("counters")[0])
    print(case, abs_mode, {k: ("%.1e" % v if isinstance(v, float) else v) for k, v in w.items()})


# ---- quirk Q8 (Particles.cpp:240-244): original particle 0 is the strict maximum along an axis -> the sequential
# `if (x<min) .. else if (x>max)` loop decides which later particles were ever compared against the maximum ----
# The parallel reduction (min over all, max over index >= 1) equals that loop unless particle 1 is the runner-up as well
# (then it too is a "record low" never tested against the maximum): only that case takes the sequential replay kernels.
@pytest.mark.parametrize("runner_up", [False, True])
@pytest.mark.parametrize("axis", [0, 1])
def test_q8_bounding_box_replay(axis, runner_up):
    ic = IC.fluid_block(40, jitter=0.05)
    key = ("x", "y")[axis]
    order = np.argsort(ic[key])
    top, second = int(order[-1]), int(order[-2])
    perm = np.arange(len(ic["x"]))
    perm[0], perm[top] = top, 0  # the extreme particle becomes original particle 0
    if runner_up:
        src = int(np.nonzero(perm == second)[0][0])
        perm[1], perm[src] = perm[src], perm[1]  # .. and the second largest original particle 1
    for k in ("x", "y", "vx", "vy", "m", "u"):
        ic[k] = np.ascontiguousarray(ic[k][perm])
    assert ic[key][0] == ic[key].max()
    if runner_up:
        assert ic[key][1] == np.sort(ic[key])[-2]
    orc = Oracle(orc_config("fb2d", ic["h"], ic["gamma"], None, abs_mode=1), ic)
    gpu = capi.MfvGpu(capi.make_config("fb2d", ic["h"], ic["gamma"], None, abs_mode=capi.ABS_FABS, debug_capture=1))
    gpu.upload(ic)
    dt_o = orc.step(stop_after=1)
    dt_g = gpu.prepare()
    cg, sg, bg = gpu.grid()
    co, so, bo = orc.grid()
    # the replayed maximum is NOT the true maximum: x[0] never reaches the `else if`
    assert bo[2 + axis] < ic[key].max()
    assert np.array_equal(cg, co) and np.array_equal(sg, so) and np.array_equal(bg, bo), (bg, bo)
    assert abs(dt_g - dt_o) <= 1e-12 * dt_o
    # particle 0 lies outside the box the reference builds: it indexes a cell out of range there (UB); here it is
    # clamped into the last cell and flagged.  Everything that does not involve it must still agree.
    flags = gpu.error_flags()
    assert flags & ~capi.F_OUT_OF_GRID == 0
    cell_g, cell_o = gpu.fetch("cell"), orc.fetch("cell")
    assert np.array_equal(cell_g[2:], cell_o[2:])
    # second step: the box comes from the update kernel's fused reduction (not the stand-alone pass) -> replay again
    orc2 = Oracle(orc.cfg, ic)
    gpu.advance(dt_o)
    orc2.step(dt_fixed=dt_o)
    dt_o2 = orc2.step(stop_after=1)
    gpu.prepare()
    cg, sg, bg = gpu.grid()
    co, so, bo = orc2.grid()
    if not np.isnan(dt_o2):
        assert np.array_equal(cg, co) and np.allclose(bg, bo, rtol=1e-12), (bg, bo)


# ---- a neighbour list cut at capacity (reference: exit(1), Particles.cpp:348-352): the device must flag it and every
# later pass must stay inside its arrays (face map of a pair whose other side was cut) ----
@pytest.mark.parametrize("case,cap", [("kh_random_50", 50), ("sedov_21", 32)])
def test_capacity_cut_lists_flag_and_stay_in_bounds(case, cap):
    ic, orc, gpu = parity.make_pair(case, capi.ABS_FABS, max_ni=cap)
    gpu.step()
    assert gpu.error_flags() & capi.F_MAX_INTERACTIONS
    noi = gpu.fetch("noi") + gpu.fetch("noiGhosts")
    assert noi.max() <= cap and noi.min() >= 0
    # the lists that were NOT cut are still the reference's
    orc.step(stop_after=1)
    o_noi = orc.fetch("noi") + (orc.fetch("noiGhosts") if ic["periodic"] else 0)
    whole = o_noi <= cap
    assert whole.any() and (~whole).any()
    assert np.array_equal(noi[whole], o_noi[whole])
    # a context that has flagged keeps working memory-safely: another step does not fault
    gpu.step()
    gpu.synchronize()


# ---- Particles::checkFluxSymmetry (Particles.cpp:2888-2976) as a device check of the slot -> face map ----
@pytest.mark.parametrize("case", ["kh_random_50", "kh_lattice_100", "sedov_21", "fb_jitter_60"])
def test_flux_symmetry_device_check(case):
    ic, orc, gpu = parity.make_pair(case, capi.ABS_FABS)
    for _ in range(2):
        gpu.step()
    before = int(gpu.fetch("counters")[0])  # one-sided seam pairs found by the searches so far (accumulates)
    gpu.prepare()
    slots, bad, two_sided, one_sided = [int(v) for v in gpu.fetch("flux_symmetry")]
    noi = gpu.fetch("noi").astype(np.int64).sum() + gpu.fetch("noiGhosts").astype(np.int64).sum()
    nfaces = int(gpu.fetch("num_faces")[0])
    assert bad == 0, "slot -> face map inconsistent for %d slots" % bad
    assert slots == noi, "every list slot maps to a face"
    assert two_sided + one_sided == nfaces and slots == 2 * two_sided + one_sided
    # faces only one endpoint uses = the one-sided periodic pairs the search counted (quirk Q9); none without periodicity
    assert one_sided == int(gpu.fetch("counters")[0]) - before
    print(case, "slots", slots, "faces", nfaces, "one-sided", one_sided)
