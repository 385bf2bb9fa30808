"""CPU: the headless diagnostics (meshlesshydro_b200/diagnostics.py) on oracle runs -- what conservationPlotter.py /
PlotSedov.py of the reference would draw."""
import numpy as np

from meshlesshydro_b200 import diagnostics as DG, ic as IC
from cpu_oracles import Oracle, make_config


def test_radial_profile_and_shock_radius_on_a_synthetic_shell():
    rng = np.random.default_rng(3)
    p = rng.uniform(-0.5, 0.5, (20000, 3))
    r = np.sqrt((p ** 2).sum(axis=1))
    rho = 1.0 + 3.0 * np.exp(-((r - 0.3) / 0.02) ** 2)  # a dense shell at r = 0.3
    rc, mean, cnt = DG.radial_profile(p[:, 0], p[:, 1], p[:, 2], rho, nbins=25, rmax=0.5)
    assert cnt.sum() == int((r <= 0.5).sum()) and cnt[2:].min() > 0
    assert abs(DG.sedov_shock_radius(p[:, 0], p[:, 1], p[:, 2], rho, nbins=25, rmax=0.5) - 0.3) <= 0.02
    # R_s = xi0 (E t^2 / rho0)^(1/5): doubling t scales R_s by 2^(2/5)
    r1, r2 = DG.sedov_shock_radius_analytic(1.0, 1.0, 0.05), DG.sedov_shock_radius_analytic(1.0, 1.0, 0.1)
    assert abs(r2 / r1 - 2.0 ** 0.4) <= 1e-12 and abs(r1 - 1.1527 * (0.0025) ** 0.2) <= 1e-12


def test_kh_oracle_run_conserves_and_keeps_the_seeded_mode():
    ic = IC.kelvin_helmholtz(32, lattice=True, jitter=0.2)  # (the uniform-random IC of this size goes NaN in the reference algorithm)
    orc = Oracle(make_config("kh2d", ic["h"], ic["gamma"], ic.get("box"), abs_mode=1), ic)
    a0 = DG.kh_mode_amplitude(ic["x"], ic["vy"], ic["m"])
    assert abs(a0 - 0.01) <= 1e-3  # vy = 0.01 sin(4 pi x), generateIC.py:42-43
    series = [orc.sums()]
    for _ in range(20):
        orc.step()
        series.append(orc.sums())
    d = DG.conservation_drift(series)
    assert d["mass"] <= 1e-12 and d["energy"] <= 1e-11 and d["momentum"] <= 1e-11, d
    a1 = DG.kh_mode_amplitude(orc.fetch("x"), orc.fetch("vy"), orc.fetch("m"))
    assert 0.2 * a0 < a1 < 5.0 * a0
