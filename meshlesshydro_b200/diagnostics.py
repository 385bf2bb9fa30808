"""Headless run diagnostics (numpy only) -- what the reference's plotting scripts draw, as numbers:

* conservation curves: /root/reference/testcases/kelvin-helmholtz/conservationPlotter.py reads the per-snapshot sums
  that Particles::dump2file writes (Particles.cpp:3011-3017) and plots them against time;
* Kelvin-Helmholtz: amplitude of the seeded vy = 0.01 sin(4 pi x) mode (generateIC.py:42-43), the usual growth measure;
* Sedov: radial density profile and shock radius against the self-similar solution
  (/root/reference/testcases/sedov/PlotSedov.py:17-211 plots rho(r) over the analytical curve).

matplotlib is not available in this image, so these return arrays/scalars for tests and logs.
"""
import numpy as np

# Sedov-Taylor: R_s = xi0 (E t^2 / rho0)^(1/5) in 3D; xi0 depends on gamma only (tabulated values of the
# self-similar solution: 1.1527 for gamma = 5/3, 1.0328 for gamma = 7/5)
_XI0 = {round(5.0 / 3.0, 6): 1.1527, round(1.4, 6): 1.0328}


def sedov_shock_radius_analytic(E, rho0, t, gamma=5.0 / 3.0):
    xi0 = _XI0.get(round(gamma, 6))
    if xi0 is None:
        raise ValueError("no tabulated xi0 for gamma=%r" % gamma)
    return xi0 * (E * t * t / rho0) ** 0.2


def radial_profile(x, y, z, q, nbins=24, rmax=None):
    """mean of q in spherical shells: (bin centres, means, counts)"""
    r = np.sqrt(x * x + y * y + z * z)
    rmax = float(r.max()) if rmax is None else rmax
    edges = np.linspace(0.0, rmax, nbins + 1)
    inside = r <= rmax  # the corners of a cubic box lie beyond the largest shell
    which = np.clip(np.digitize(r[inside], edges) - 1, 0, nbins - 1)
    cnt = np.bincount(which, minlength=nbins)
    tot = np.bincount(which, weights=np.asarray(q)[inside], minlength=nbins)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.where(cnt > 0, tot / np.maximum(cnt, 1), np.nan)
    return 0.5 * (edges[1:] + edges[:-1]), mean, cnt


def sedov_shock_radius(x, y, z, rho, nbins=24, rmax=0.5):
    """radius of the densest shell"""
    rc, mean, cnt = radial_profile(x, y, z, rho, nbins, rmax)
    ok = cnt > 0
    return float(rc[ok][np.nanargmax(mean[ok])])


def kh_mode_amplitude(x, vy, m, k=2):
    """mass-weighted amplitude of the sin(2 pi k x) component of vy (k = 2 is the seeded mode)"""
    s = np.sum(m * vy * np.sin(2.0 * np.pi * k * x))
    c = np.sum(m * vy * np.cos(2.0 * np.pi * k * x))
    return 2.0 * float(np.hypot(s, c) / np.sum(m))


def conservation_drift(series):
    """series: list of 6-vectors [volume, mass, energy, px, py, pz] (mlh_sums order) -> dict of the largest drift
    relative to the initial mass / energy / momentum scale sqrt(2 M E)"""
    a = np.asarray(series, dtype=np.float64)
    M, E = a[0, 1], a[0, 2]
    pscale = np.sqrt(2.0 * M * E)
    return {"mass": float(np.max(np.abs(a[:, 1] - M)) / M), "energy": float(np.max(np.abs(a[:, 2] - E)) / E),
            "momentum": float(np.max(np.abs(a[:, 3:6] - a[0, 3:6])) / pscale)}
