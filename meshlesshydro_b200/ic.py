"""Synthetic initial conditions of the reference's test cases (no HDF5 needed).

Formulas restate the reference's IC generators (they cannot run here: h5py, seagen and
matplotlib are absent) so that every arm -- GPU path, CPU oracle, reference-source build --
is fed identical particles:

* Kelvin-Helmholtz 2D: /root/reference/testcases/kelvin-helmholtz/generateIC.py:11-59,78-93,109-111
* Sedov 3D:            /root/reference/testcases/sedov/initial_sedov.py:8-26,33-40,64-71
  (SEAGen glass replaced by a seeded, jittered cubic lattice -- SURVEY.md section 8d)
* fluid block 2D:      /root/reference/testcases/fluid-block/generateIC.py:26-57 (``-t``)

All functions return a dict of float64 numpy arrays ``x y [z] vx vy [vz] m u`` plus the run
parameters ``dim periodic box h gamma`` the reference would read from config.info.
"""
import numpy as np

GAMMA = 5.0 / 3.0
SEED = 6102003  # generateIC.py:91


def _kh_vx(y):
    v1, v2, Dy = 0.5, -0.5, 0.025
    Dv = (v1 - v2) / 2.0
    out = np.empty_like(y)
    m1 = y < 0.25
    out[m1] = v1 - Dv * np.exp((y[m1] - 0.25) / Dy)
    m2 = (0.25 <= y) & (y < 0.5)
    out[m2] = v2 + Dv * np.exp((0.25 - y[m2]) / Dy)
    m3 = (0.5 <= y) & (y <= 0.75)
    out[m3] = v2 + Dv * np.exp((y[m3] - 0.75) / Dy)
    m4 = 0.75 < y
    out[m4] = v1 - Dv * np.exp((0.75 - y[m4]) / Dy)
    return out


def _kh_rho(y):
    rho1, rho2, Dy = 1.0, 2.0, 0.025
    Drho = (rho1 - rho2) / 2.0
    out = np.empty_like(y)
    m1 = y < 0.25
    out[m1] = rho1 - Drho * np.exp((y[m1] - 0.25) / Dy)
    m2 = (0.25 <= y) & (y < 0.5)
    out[m2] = rho2 + Drho * np.exp((0.25 - y[m2]) / Dy)
    m3 = (0.5 <= y) & (y <= 0.75)
    out[m3] = rho2 + Drho * np.exp((y[m3] - 0.75) / Dy)
    m4 = 0.75 < y
    out[m4] = rho1 - Drho * np.exp((0.75 - y[m4]) / Dy)
    return out


def kelvin_helmholtz(n_side, lattice=True, h_over_dx=4.0, jitter=0.0):
    """KH 2D, N = n_side**2, periodic box [0,1]^2, h = h_over_dx/n_side.

    lattice=True  -> ``generateIC.py -g`` (x outer, y inner, linspace endpoint=False)
    lattice=False -> ``default_rng(6102003).random((N, 2))``
    jitter > 0    -> lattice displaced by uniform +-jitter*dx (tie-free periodic seam)
    """
    N = n_side * n_side
    if lattice:
        ax = np.linspace(0.0, 1.0, n_side, endpoint=False)
        x = np.repeat(ax, n_side)
        y = np.tile(ax, n_side)
        if jitter > 0.0:
            rng = np.random.Generator(np.random.PCG64(SEED))
            d = (rng.random((N, 2)) * 2.0 - 1.0) * (jitter / n_side)
            x = np.mod(x + d[:, 0], 1.0)
            y = np.mod(y + d[:, 1], 1.0)
    else:
        pos = np.random.default_rng(SEED).random(size=(N, 2))
        x = pos[:, 0].copy()
        y = pos[:, 1].copy()
    P = 5.0 / 2.0
    rho = _kh_rho(y)
    return dict(
        dim=2, periodic=1, box=np.array([0.0, 0.0, 1.0, 1.0]), h=h_over_dx / n_side, gamma=GAMMA,
        x=x, y=y, vx=_kh_vx(y), vy=0.01 * np.sin(4.0 * np.pi * x),
        m=rho / N, u=P / ((GAMMA - 1.0) * rho),
    )


def _sedov_w(r, sml):
    """initial_sedov.py:8-26: IC's own spline with f = 8/(pi s^3), branches q>1, q>.5, else."""
    q = r / sml
    f = 8.0 / np.pi / (sml * sml * sml)
    w = np.where(q > 1.0, 0.0, np.where(q > 0.5, 2.0 * f * (1.0 - q) ** 3, f * (6.0 * q ** 3 - 6.0 * q ** 2 + 1.0)))
    return w


def sedov(n_side, jitter=0.05):
    """Sedov blast 3D, N = n_side**3 in [-.5,.5]^3, non-periodic, h = 0.07*31/n_side.

    Jittered cubic lattice (uniform +-jitter*dx, PCG64 seed 6102003); m = 1/N, v = 0,
    u = max(W(r; 2 sml), 1e-6) with sml = 0.041833*61/n_side (initial_sedov.py:33-34,40,64-71).
    """
    N = n_side ** 3
    dx = 1.0 / n_side
    ax = (np.arange(n_side) + 0.5) * dx - 0.5
    x = np.repeat(ax, n_side * n_side)
    y = np.tile(np.repeat(ax, n_side), n_side)
    z = np.tile(ax, n_side * n_side)
    rng = np.random.Generator(np.random.PCG64(SEED))
    d = (rng.random((N, 3)) * 2.0 - 1.0) * (jitter * dx)
    x = x + d[:, 0]
    y = y + d[:, 1]
    z = z + d[:, 2]
    sml = 0.041833 * 61.0 / n_side
    r = np.sqrt(x * x + y * y + z * z)
    u = np.maximum(_sedov_w(r, 2.0 * sml), 1e-6)
    zero = np.zeros(N)
    return dict(
        dim=3, periodic=0, box=None, h=0.07 * 31.0 / n_side, gamma=GAMMA,
        x=x, y=y, z=z, vx=zero.copy(), vy=zero.copy(), vz=zero.copy(),
        m=np.full(N, 1.0 / N), u=u,
    )


def warm_blob_3d(n_side):
    """Smooth 3D test state on the Sedov particle arrangement: u = 1 + 0.2 exp(-r^2/0.05), a gentle shear flow.
    Not one of the reference's cases -- for variants that the blast wave turns into NaN in the reference itself
    (FIRST_ORDER_QUAD_POINT 0 extrapolates the neighbour's state over ~the full separation)."""
    ic = sedov(n_side)
    r2 = ic["x"] ** 2 + ic["y"] ** 2 + ic["z"] ** 2
    ic["u"] = 1.0 + 0.2 * np.exp(-r2 / 0.05)
    ic["vx"] = 0.1 * np.sin(2.0 * np.pi * ic["y"])
    ic["vy"] = 0.05 * np.cos(2.0 * np.pi * ic["z"])
    ic["vz"] = 0.05 * np.sin(2.0 * np.pi * ic["x"])
    return ic


def fluid_block(n_side, h_over_dx=3.5, jitter=0.0):
    """fluid-block 2D (``generateIC.py -t``): lattice in [-.5,.5)^2, v = 0, rho = 1, u = 1."""
    N = n_side * n_side
    ax = np.linspace(-0.5, 0.5, n_side, endpoint=False)
    x = np.repeat(ax, n_side)
    y = np.tile(ax, n_side)
    if jitter > 0.0:
        rng = np.random.Generator(np.random.PCG64(SEED))
        d = (rng.random((N, 2)) * 2.0 - 1.0) * (jitter / n_side)
        x = x + d[:, 0]
        y = y + d[:, 1]
    zero = np.zeros(N)
    return dict(
        dim=2, periodic=0, box=None, h=h_over_dx / n_side, gamma=GAMMA,
        x=x, y=y, vx=zero.copy(), vy=zero.copy(), m=np.full(N, 1.0 / N), u=np.ones(N),
    )
