"""GPU (-m gpu): the CUDA path through the C ABI against the committed golden vectors, i.e. against the
outputs of the reference's OWN sources (oracle/_ref) -- no oracle code involved at run time.
Index work bit-exact (cells, ordered neighbour lists); floating point <= 1e-10 (rule: parity.close)."""
import numpy as np
import pytest

import golden_util as G
import parity
from meshlesshydro_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", G.NAMES)
def test_gpu_matches_reference_golden(name):
    g, ic = G.load(name)
    D = g["dim"]
    cap = 160
    cfg = capi.make_config(g["preset"], g["h"], g["gamma"], ic["box"], abs_mode=g["abs_mode"], debug_capture=1,
                           max_interactions=cap)
    gpu = capi.MfvGpu(cfg)
    gpu.upload(ic)
    dt = gpu.prepare()
    assert gpu.error_flags() == 0
    assert abs(dt - g["dt_cfl"]) <= 1e-12 * g["dt_cfl"]
    cells, cs, _ = gpu.grid()
    assert np.array_equal(cells, g["cells"])
    assert np.array_equal(cs, g["cell_size"])
    assert np.array_equal(gpu.fetch("cell"), g["cell"])
    noi = gpu.fetch("noi")
    assert np.array_equal(noi, g["noi"])
    assert np.array_equal(G.compact(gpu.fetch("nnl"), noi, cap), g["nnl"])
    if g["periodic"]:
        ng = gpu.fetch("noiGhosts")
        assert np.array_equal(ng, g["noiGhosts"])
        assert np.array_equal(G.compact(gpu.fetch("nnlGhosts"), ng, cap), g["nnlGhostParents"])
    for k in ["omega", "rho", "P", "gradPre", "rhoGrad", "vxGrad", "vyGrad", "PGrad"] + (["vzGrad"] if D == 3 else []):
        parity.close(gpu.fetch(k), g[k], what=k)
    gpu.advance(g["dt_cfl"])
    assert gpu.error_flags() & ~capi.F_NEG_GHOST_PRESSURE == 0
    for k in ("mF", "eF", "vF"):
        parity.close(gpu.fetch(k), g["out_" + k], what=k)
    st = gpu.download_state()
    for k in ["x", "y", "vx", "vy", "m", "u"] + (["z", "vz"] if D == 3 else []):
        parity.close(st[k], g["out_" + k], what="out_" + k)
    s1 = gpu.sums()
    assert np.allclose(s1[1:], g["sums1"][1:], rtol=1e-12, atol=1e-15)
    gpu.close()
