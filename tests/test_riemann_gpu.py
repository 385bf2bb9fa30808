"""GPU: the stand-alone face solver (mlh_riemann_faces = the reference's Riemann class, Riemann.cpp:7-229) against
the oracle's orc_face_flux on random faces: smooth states, strong shocks, strong rarefactions, identical states."""
import ctypes as C
import os

import numpy as np
import pytest

from meshlesshydro_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_fluxes(dim, mfm, gamma, WR, WL, vF, A):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    dp = C.POINTER(C.c_double)
    lib.orc_face_flux.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp, dp, dp, dp]
    F = np.zeros_like(WR)
    for k in range(len(WR)):
        wr, wl = WR[k].copy(), WL[k].copy()  # modified in place like the reference
        lib.orc_face_flux(dim, mfm, gamma, wr.ctypes.data_as(dp), wl.ctypes.data_as(dp), vF[k].ctypes.data_as(dp),
                          A[k].ctypes.data_as(dp), F[k].ctypes.data_as(dp))
    return F


def _faces(dim, n, seed):
    rng = np.random.default_rng(seed)
    nw = dim + 2
    WL = np.empty((n, nw))
    WR = np.empty((n, nw))
    kind = rng.integers(0, 4, n)
    for W in (WL, WR):
        W[:, 0] = rng.uniform(0.2, 3.0, n)
        W[:, 1] = rng.uniform(0.2, 3.0, n)
        W[:, 2:] = rng.normal(0, 0.4, (n, dim))
    strong = kind == 1                       # strong shocks: pressure ratios up to 1e4
    WL[strong, 1] *= 10 ** rng.uniform(0, 4, strong.sum())
    rare = kind == 2                         # receding flows: strong rarefactions (no vacuum)
    WL[rare, 2] -= 1.0
    WR[rare, 2] += 1.0
    same = kind == 3                         # identical states: f(P) = 0 exactly, zero iterations
    WR[same] = WL[same]
    A = rng.normal(0, 1, (n, dim)) * 10 ** rng.uniform(-4, 0, (n, 1))
    A[:, 0] = np.abs(A[:, 0]) + 1e-6         # away from the -x singularity of the 3D rotation (Helper.cpp:48-77)
    vF = rng.normal(0, 0.3, (n, dim))
    return WR, WL, vF, A


@pytest.mark.parametrize("dim,gamma", [(2, 5.0 / 3.0), (3, 5.0 / 3.0), (2, 1.4), (3, 1.3)])
def test_face_fluxes_match_oracle(dim, gamma):
    WR, WL, vF, A = _faces(dim, 4000, seed=dim * 10 + int(gamma * 10))
    cfg = capi.make_config("fb2d" if dim == 2 else "sedov3d", 1.0, gamma)
    gpu = capi.MfvGpu(cfg)
    F = gpu.riemann_faces(WR, WL, vF, A)
    ref = _oracle_fluxes(dim, 0, gamma, WR, WL, vF, A)
    scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-300
    err = np.abs(F - ref) / scale
    assert err.max() <= 1e-10, (err.max(), np.unravel_index(err.argmax(), err.shape))
    # identical left and right states: the flux is the analytic flux of that state, no iteration involved
    gpu.close()


def test_inputs_untouched_and_antisymmetry():
    WR, WL, vF, A = _faces(2, 500, seed=7)
    cfg = capi.make_config("fb2d", 1.0, 5.0 / 3.0)
    gpu = capi.MfvGpu(cfg)
    wr0, wl0 = WR.copy(), WL.copy()
    F = gpu.riemann_faces(WR, WL, vF, A)
    assert np.array_equal(WR, wr0) and np.array_equal(WL, wl0)
    # the same face seen from the other side (states swapped, face reversed): flux flips sign (to round-off)
    G = gpu.riemann_faces(WL, WR, vF, -A)
    scale = np.abs(F).max(axis=1, keepdims=True) + 1e-300
    assert (np.abs(F + G) / scale).max() <= 1e-9
    gpu.close()
