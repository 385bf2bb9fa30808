"""CPU: the C-ABI library loads and exports every symbol include/mlh_gpu.h declares; host-only entry
points behave; without a CUDA device the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from meshlesshydro_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mlh_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(mlh_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_symbols_exported():
    lib = C.CDLL(capi.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 30
    missing = [s for s in decl if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes binding covers the whole header
    assert sorted(capi.EXPORTED_SYMBOLS) == decl


def test_abi_version_and_defaults():
    lib = capi.load_library()
    assert lib.mlh_abi_version() == 2
    cfg = capi.default_config()
    # demonstrator/include/parameter.h values
    assert (cfg.dim, cfg.periodic, cfg.slope_limiting, cfg.pairwise_limiter, cfg.move_particles) == (2, 1, 1, 1, 1)
    assert cfg.cfl == pytest.approx(0.2) and cfg.beta == pytest.approx(4.0)
    assert cfg.nranks == 1


def test_config_struct_layout_matches_header():
    """sizeof(mlh_config) as the C compiler sees it == the ctypes mirror (guards silent ABI drift)."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "mlh_gpu.h"\nint main(){printf("%zu %zu %zu", sizeof(mlh_config), ' \
          '__builtin_offsetof(mlh_config, cfl), __builtin_offsetof(mlh_config, capacity));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        size, off_cfl, off_cap = map(int, subprocess.check_output([os.path.join(d, "t")]).split())
    assert size == C.sizeof(capi.MlhConfig)
    assert off_cfl == capi.MlhConfig.cfl.offset
    assert off_cap == capi.MlhConfig.capacity.offset


def test_slab_range_partitions_layers():
    lib = capi.load_library()
    for layers, nranks in [(500, 8), (117, 8), (25, 2), (7, 7), (28, 3)]:
        covered = []
        for r in range(nranks):
            lo, hi = C.c_int(), C.c_int()
            assert lib.mlh_slab_range(layers, nranks, r, C.byref(lo), C.byref(hi)) == 0
            assert hi.value > lo.value
            covered += list(range(lo.value, hi.value))
            assert abs((hi.value - lo.value) - layers / nranks) < 1.0
        assert covered == list(range(layers))
    lo, hi = C.c_int(), C.c_int()
    assert lib.mlh_slab_range(10, 4, 4, C.byref(lo), C.byref(hi)) < 0


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    cfg = capi.make_config("kh2d", 0.04, 5.0 / 3.0, [0, 0, 1, 1])
    with pytest.raises(capi.MlhError, match="no CUDA device|no CPU fallback"):
        capi.MfvGpu(cfg)


def test_invalid_config_rejected():
    lib = capi.load_library()
    cfg = capi.default_config()
    cfg.dim = 4
    ctx = C.c_void_p()
    assert lib.mlh_create(C.byref(cfg), C.byref(ctx)) == -1
    assert b"dim" in lib.mlh_last_error(None)
