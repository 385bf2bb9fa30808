"""CPU: diagnostics.SedovTaylor (the similarity solution PlotSedov.py:17-175 draws over the reference's snapshots)
against the published energy constants alpha = E t^2 / (rho0 R_s^(nu+2)) for gamma = 1.4 (Kamm & Timmes 2007, Table 1:
planar 0.5387 per side, cylindrical 0.9841, spherical 0.8511) and against its own conservation laws."""
import numpy as np
import pytest

from meshlesshydro_b200 import diagnostics as DG


@pytest.mark.parametrize("gamma,nu,alpha", [(1.4, 3, 0.851072), (1.4, 2, 0.984074), (1.4, 1, 2 * 0.538743)])
def test_energy_constant_matches_published_values(gamma, nu, alpha):
    s = DG.SedovTaylor(1.0, 1.0, gamma, nu)
    assert abs(s.xi0 ** -(nu + 2) - alpha) <= 2e-5 * alpha


def test_gamma_five_thirds_spherical_front():
    s = DG.SedovTaylor(energy=1.0, rho0=1.0, gamma=5.0 / 3.0, nu=3)  # the reference's Sedov test case
    assert abs(s.xi0 - 1.15167) <= 2e-5
    assert abs(s.shock_radius(0.01) - 1.15167 * 0.01 ** 0.4) <= 1e-5
    assert abs(s.post_shock_density - 4.0) <= 1e-12


@pytest.mark.parametrize("gamma,nu", [(1.4, 3), (5.0 / 3.0, 3), (5.0 / 3.0, 2), (1.4, 1)])
def test_swept_up_mass_and_energy_are_conserved(gamma, nu):
    E, rho0, t = 2.5, 0.7, 0.3
    s = DG.SedovTaylor(E, rho0, gamma, nu)
    Rs = s.shock_radius(t)
    r = np.linspace(1e-5 * Rs, Rs * (1 - 1e-10), 400001)
    geom = {1: 2.0, 2: 2.0 * np.pi, 3: 4.0 * np.pi}[nu]
    dV = geom * r ** (nu - 1)
    rho, v, P = s.density(r, t), s.velocity(r, t), s.pressure(r, t)
    mass = np.trapezoid(rho * dV, r)
    assert abs(mass / (rho0 * geom * Rs ** nu / nu) - 1.0) <= 1e-5   # all the swept-up gas sits behind the front
    energy = np.trapezoid((0.5 * rho * v * v + P / (gamma - 1.0)) * dV, r)
    assert abs(energy / E - 1.0) <= 1e-4
    assert abs(rho[-1] / rho0 - (gamma + 1.0) / (gamma - 1.0)) <= 1e-5  # strong-shock jump
    assert abs(v[-1] - 2.0 / (gamma + 1.0) * s.shock_speed(t)) <= 1e-5 * s.shock_speed(t)


def test_sedov_front_finds_a_synthetic_shell():
    rng = np.random.default_rng(3)
    n = 200000
    pts = rng.uniform(-0.5, 0.5, (n, 3))
    r = np.linalg.norm(pts, axis=1)
    rho = 1.0 + 3.0 * np.exp(-((r - 0.2) / 0.01) ** 2)  # a thin dense shell at r = 0.2
    r_peak, rho_peak, r_half = DG.sedov_front(pts[:, 0], pts[:, 1], pts[:, 2], rho, rho0=1.0, nbins=50, rmax=0.5)
    assert abs(r_peak - 0.2) <= 0.011 and rho_peak > 2.0 and r_half >= r_peak
