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


class SedovTaylor:
    """Self-similar point-blast solution in a uniform medium (Sedov 1959; closed parametric form as in Book 1994,
    "The Sedov self-similar point blast solutions in nonuniform media", with w = 0), the curve the reference's
    PlotSedov.py draws over its snapshots (/root/reference/testcases/sedov/PlotSedov.py:17-175).

    With V the similarity variable running from V_min = 2/((nu+2) gamma) (centre; shown here scaled by (nu+2)/2 as f)
    to the shock, the profiles are products of powers of three linear factors of f.  R_s(t) = c (E t^2 / rho0)^(1/(nu+2))
    where c follows from energy conservation: E = int (rho v^2/2 + P/(gamma-1)) dV.
    """

    def __init__(self, energy=1.0, rho0=1.0, gamma=5.0 / 3.0, nu=3, samples=200001):
        g, n = float(gamma), int(nu)
        if n not in (1, 2, 3):
            raise ValueError("nu must be 1, 2 or 3")
        self.E, self.rho0, self.gamma, self.nu = float(energy), float(rho0), g, n
        # exponents of the three factors (uniform medium)
        w1 = (3.0 * n - 2.0 + g * (2.0 - n)) / (g + 1.0)
        w2 = (2.0 * (g - 1.0) + n) / g
        w3 = n * (2.0 - g)
        a0 = 1.0 / (n * g - n + 2.0)
        a2 = (g - 1.0) / (g * w2)
        a3 = n / (g * w2)
        a5 = 2.0 * n / w3
        a6 = 2.0 / (n + 2.0)
        a1 = a2 + (g + 1.0) * a0 - a6
        a4 = a1 * n * (n + 2.0) / w3
        a8 = n * a6
        k5 = 2.0 / (g - 1.0)
        k6 = 0.5 * (g + 1.0)
        k1 = k5 * g
        k2 = k6 / g
        k3 = (n * g - n + 2.0) / (w1 * k6)
        k4 = (n + 2.0) * a0 * k6
        # from the centre (f -> k2, where r/R_s ~ (f - k2)^a2 with a2 ~ 0.1) to the shock (f = 1): the distance from k2
        # is sampled logarithmically down to 1e-60, i.e. r/R_s < 1e-3 -- sampling f itself left the inner quarter of
        # the radius, where the pressure stays finite, out of the energy integral (xi0 0.24 % too large)
        delta = np.logspace(-60.0, np.log10(1.0 - k2), int(samples))
        f = k2 + delta
        A, B, C = k1 * delta, k3 * (k4 - f), k5 * (k6 - f)
        eta = f ** (-a6) * A ** a2 * B ** (-a1)          # r / R_s
        dens = A ** a3 * B ** a4 * C ** (-a5)            # rho / rho_shock
        pres = f ** a8 * B ** (a4 - 2.0 * a1) * C ** (1.0 - a5)  # P / P_shock
        vel = eta * f                                    # v / v_shock
        order = np.argsort(eta)
        self._eta, self._d, self._p, self._v = eta[order], dens[order], pres[order], vel[order]
        # normalisation: with the shock values rho_s = (g+1)/(g-1) rho0, v_s = 2/(g+1) D, P_s = 2/(g+1) rho0 D^2 and
        # D = a6 R_s / t the energy integral gives  E t^2 / (rho0 R_s^(nu+2)) = alpha
        geom = {1: 2.0, 2: 2.0 * np.pi, 3: 4.0 * np.pi}[n]
        integrand = self._eta ** (n - 1) * (self._d * self._v ** 2 + self._p)
        integral = np.sum(0.5 * (integrand[1:] + integrand[:-1]) * np.diff(self._eta))
        alpha = integral * 8.0 * geom / ((g * g - 1.0) * (n + 2.0) ** 2)
        self.xi0 = alpha ** (-1.0 / (n + 2.0))

    def shock_radius(self, t):
        return self.xi0 * (self.E * np.asarray(t, dtype=np.float64) ** 2 / self.rho0) ** (1.0 / (self.nu + 2.0))

    def shock_speed(self, t):
        return 2.0 / (self.nu + 2.0) * self.shock_radius(t) / t

    @property
    def post_shock_density(self):
        return (self.gamma + 1.0) / (self.gamma - 1.0) * self.rho0

    def density(self, r, t):
        eta = np.asarray(r, dtype=np.float64) / self.shock_radius(t)
        inside = np.interp(eta, self._eta, self._d, left=0.0) * self.post_shock_density
        return np.where(eta <= 1.0, inside, self.rho0)

    def pressure(self, r, t):
        eta = np.asarray(r, dtype=np.float64) / self.shock_radius(t)
        ps = 2.0 / (self.gamma + 1.0) * self.rho0 * self.shock_speed(t) ** 2
        return np.where(eta <= 1.0, np.interp(eta, self._eta, self._p, left=self._p[0]) * ps, 0.0)

    def velocity(self, r, t):
        eta = np.asarray(r, dtype=np.float64) / self.shock_radius(t)
        vs = 2.0 / (self.gamma + 1.0) * self.shock_speed(t)
        return np.where(eta <= 1.0, np.interp(eta, self._eta, self._v, left=0.0) * vs, 0.0)


def sedov_front(x, y, z, rho, rho0=1.0, nbins=40, rmax=0.5):
    """(radius of the densest shell, its mean density, radius where the shell density last exceeds the mean of
    background and peak -- the half-rise point of the front seen from outside)"""
    rc, mean, cnt = radial_profile(x, y, z, rho, nbins, rmax)
    ok = cnt > 0
    rc, mean = rc[ok], mean[ok]
    kmax = int(np.nanargmax(mean))
    half = 0.5 * (mean[kmax] + rho0)
    above = np.nonzero(mean >= half)[0]
    return float(rc[kmax]), float(mean[kmax]), float(rc[above[-1]])
