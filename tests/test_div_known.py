"""CPU: the division K3/K3b use for r/(h/2) and W/omega_i (mlh_div_known, csrc/mlh_internal.cuh: multiplication by the
correctly rounded reciprocal + two FMA corrections, Markstein) returns the IEEE quotient bit for bit -- the reference
divides (Particles.cpp:10, :1204, :1247), and limiter decisions downstream sit on knife edges for lattice initial
conditions, so 'almost' would not do.  The GPU side is covered by the parity tests and tools/bitwise_ab.py."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _lib(tmp_path):
    so = os.path.join(str(tmp_path), "libdivknown.so")
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "c", "div_known.c"), "-lm"], check=True)
    lib = ctypes.CDLL(so)
    lib.div_known_mismatches.restype = ctypes.c_long
    lib.div_known_mismatches.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_long, ctypes.c_uint64]
    return lib


def test_division_by_known_reciprocal_is_correctly_rounded(tmp_path):
    lib = _lib(tmp_path)
    rng = np.random.default_rng(5)
    # h/2 of the shipped and benchmarked cases, awkward significands, then random divisors (h/2 and omega ranges)
    fixed = [0.001, 0.00175, 0.002, 0.0035, 0.01, 0.0125, 0.02, 0.0375, 0.05, 1.1 / 61 / 2, 0.3, 0.7, 3.0, 0.1,
             np.nextafter(2.0, 1.0), np.nextafter(1.0, 2.0), 1.0, 2.0 ** -20, 7.0e8, 1.23456789e5]
    div = np.array(fixed + list(np.exp(rng.uniform(np.log(1e-6), np.log(1e10), 600))), dtype=np.float64)
    bad = lib.div_known_mismatches(div.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(div), 20000, 12345)
    assert bad == 0, "%d quotients differ from the IEEE division" % bad
