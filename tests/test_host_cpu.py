"""CPU: host side of the drop-in (meshlesshydro_b200/host): the HDF5 subset reader/writer, the config.info parser
and the host-only parts of the C++ mirror classes.  No GPU, no /root/reference."""
import os
import struct
import subprocess

import numpy as np
import pytest

from meshlesshydro_b200 import h5lite

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "meshlesshydro_b200", "host")

SNAPSHOT_NAMES = ["time", "totalMass", "energy", "xMomentum", "yMomentum", "zMomentum", "rho", "m", "u", "x", "v",
                  "rhoGrad", "P", "noi"]  # Particles.cpp:2990-3006


def _snapshot(n=37, dim=3, seed=1):
    rng = np.random.default_rng(seed)
    d = {}
    for k in SNAPSHOT_NAMES:
        if k in ("x", "v", "rhoGrad"):
            d[k] = rng.standard_normal((n, dim))
        elif k == "noi":
            d[k] = rng.integers(0, 200, n).astype(np.int32)
        elif k in ("rho", "m", "u", "P"):
            d[k] = rng.random(n)
        else:
            d[k] = rng.random(1)
    return d


def test_h5_roundtrip_snapshot_layout(tmp_path):
    d = _snapshot()
    d["materialId"] = (np.arange(37) % 5 - 2).astype(np.int8)
    path = str(tmp_path / "snap.h5")
    h5lite.write_file(path, d)
    back = h5lite.read_file(path)
    assert sorted(back) == sorted(d)
    for k, v in d.items():
        assert back[k].shape == v.shape, k
        assert np.array_equal(back[k], v), k   # bit-exact
        assert back[k].dtype == (np.float64 if v.dtype.kind == "f" else np.int32)


def test_h5_file_structure(tmp_path):
    """The bytes follow the HDF5 spec structures libhdf5 expects (superblock v0, symbol-table root group)."""
    path = str(tmp_path / "s.h5")
    h5lite.write_file(path, _snapshot(n=5, dim=2))
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8] == 0 and b[13] == 8 and b[14] == 8           # superblock version 0, 8-byte offsets and lengths
    leaf_k, internal_k = struct.unpack_from("<HH", b, 16)
    assert (leaf_k, internal_k) == (4, 16)
    base, freesp, eof, driver = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and freesp == 2**64 - 1 and driver == 2**64 - 1 and eof == len(b)
    name_off, root_hdr, cache_type = struct.unpack_from("<QQI", b, 56)
    btree, heap = struct.unpack_from("<QQ", b, 80)
    assert cache_type == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    assert b[root_hdr] == 1                                   # version-1 object header
    mtype, msize = struct.unpack_from("<HH", b, root_hdr + 16)
    assert (mtype, msize) == (0x11, 16)                       # symbol table message
    assert struct.unpack_from("<QQ", b, root_hdr + 24) == (btree, heap)
    used = struct.unpack_from("<H", b, btree + 6)[0]
    assert used == 2                                          # 14 links -> two symbol table nodes of <= 8 entries
    heap_size, free_off, heap_data = struct.unpack_from("<QQQ", b, heap + 8)
    assert struct.unpack_from("<QQ", b, heap_data + free_off) == (1, heap_size - free_off)  # one free block, H5HL_FREE_NULL
    names = []
    for k in range(used):
        snod = struct.unpack_from("<Q", b, btree + 24 + 8 + 16 * k)[0]
        assert b[snod:snod + 4] == b"SNOD"
        for e in range(struct.unpack_from("<H", b, snod + 6)[0]):
            off, hdr = struct.unpack_from("<QQ", b, snod + 8 + 40 * e)
            end = b.index(b"\0", heap_data + off)
            names.append(b[heap_data + off:end].decode())
            assert b[hdr] == 1
    assert names == sorted(names) and set(names) == set(SNAPSHOT_NAMES)  # strcmp order, as libhdf5 keeps them


def test_h5_reads_libhdf5_file():
    """A file written by the real HDF5 library (MATLAB v7.3 = HDF5 with a 512-byte user block, ships with scipy)."""
    import scipy
    path = os.path.join(os.path.dirname(scipy.__file__), "io", "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy test data not installed")
    d = h5lite.read_file(path)
    assert list(d) == ["testdouble"]
    assert d["testdouble"].shape == (9, 1)
    assert np.allclose(d["testdouble"][:, 0], np.arange(9) * np.pi / 4, rtol=1e-15)


def test_h5_errors(tmp_path):
    p = tmp_path / "not.h5"
    p.write_bytes(b"hello world" * 10)
    with pytest.raises(h5lite.H5Error):
        h5lite.read_file(str(p))
    with pytest.raises(h5lite.H5Error):
        h5lite.read_file(str(tmp_path / "missing.h5"))


CONFIG = """; h5 file containing the initial particle distribution
initFile ../testcases/kelvin-helmholtz/khN16384.h5
outDir "out put"   ; quoted value, trailing comment
timeStep .005
timeEnd 4.
h5DumpInterval 10
periodicBoxLimits {
    lowerX 0.
    lowerY -0.25
    upperX 1.
    upperY 1.
}
someList
{
    a 1.5
    b 2.5
}
objects {
    first { name alpha }
    second { name beta }
}
kernelSize .025
gamma 1.6666666666666667
"""


def _selftest(case, *args):
    exe = os.path.join(HOST, "bin", "host_selftest_" + case)
    assert os.path.exists(exe), "run __graft_entry__.build()"
    out = subprocess.run([exe] + list(args), stdout=subprocess.PIPE, text=True, check=True).stdout
    res = {}
    for line in out.splitlines():
        k, _, v = line.partition(" ")
        res[k] = v
    return res


def test_config_parser_info_format(tmp_path):
    cfg = tmp_path / "config.info"
    cfg.write_text(CONFIG)
    r = _selftest("kh2d", str(cfg))
    assert r["initFile"] == "../testcases/kelvin-helmholtz/khN16384.h5"
    assert float(r["timeStep"]) == 0.005 and int(r["h5DumpInterval"]) == 10
    assert float(r["gamma"]) == 1.6666666666666667
    assert float(r["upperX"]) == 1.0 and float(r["nested"]) == -0.25
    assert r["missing"] == "throws"                      # boost::property_tree throws ptree_bad_path
    assert (float(r["list0"]), float(r["list1"])) == (1.5, 2.5)
    assert (r["obj0"], r["obj1"]) == ("alpha", "beta")


@pytest.mark.parametrize("case,dim", [("kh2d", 2), ("sedov3d", 3)])
def test_helper_domain_particles_host_parts(case, dim):
    r = _selftest(case)
    assert int(r["DIM"]) == dim
    inv = np.array([float(v) for v in r["inv"].split()]).reshape(dim, dim)
    A = np.array([[4., 1.], [1., 3.]]) if dim == 2 else np.array([[4., 1., .5], [1., 3., .25], [.5, .25, 2.]])
    assert np.allclose(inv, np.linalg.inv(A), rtol=1e-14)
    L = np.array([float(v) for v in r["rot"].split()]).reshape(dim, dim)
    a = np.array([0.6, 0.8]) if dim == 2 else np.array([0.36, 0.48, 0.8])
    e = np.zeros(dim)
    e[0] = 1.0
    assert np.allclose(L @ a, e, atol=1e-15)              # rotates a onto x (Helper.cpp:39-77)
    assert np.allclose(L @ L.T, np.eye(dim), atol=1e-15)
    cells = r["cells"].split()
    assert cells[:2] == ["14", "14"] and int(cells[3]) == 14 ** dim     # floor(1/0.07) = 14 cells per axis
    nb = [int(v) for v in r["nb0"].split()]
    assert len(nb) == 3 ** dim and nb.count(-1) == 3 ** dim - 2 ** dim  # corner cell: stencil clipped, no wrap
    if dim == 2:
        assert nb == [-1, -1, -1, -1, 0, 14, -1, 1, 15]    # x outer, y inner (Domain.cpp:91-100)
    lim = [float(v) for v in r["limits"].split()]
    assert lim[0] == -0.3 and lim[dim] == 0.7               # Particles::getDomainLimits
    assert r["logger"] == "42"                              # INFO line printed, DEBUG line filtered


def test_multi_rank_launcher_fails_loudly_without_gpus(tmp_path):
    """`mlh_kh2d --ranks 2` on a box without GPUs: both rank processes must end with the no-device error (exit 20) --
    no CPU fallback, and no rank left hanging in a barrier (host/src/MultiGpu.cpp)."""
    import subprocess
    import numpy as np
    from meshlesshydro_b200 import h5lite, ic as IC
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    exe = os.path.join(HOST, "bin", "mlh_kh2d")
    if not os.path.exists(exe):
        pytest.skip("host executables not built")
    ic = IC.kelvin_helmholtz(16, lattice=False)
    init = str(tmp_path / "init.h5")
    h5lite.write_initial_conditions(init, ic)
    (tmp_path / "out").mkdir()
    cfg = tmp_path / "config.info"
    cfg.write_text("initFile %s\noutDir %s\ntimeStep 0.01\ntimeEnd 0.02\nh5DumpInterval 1\nperiodicBoxLimits {\n lowerX 0\n lowerY 0\n"
                   " upperX 1\n upperY 1\n}\nkernelSize %.17g\ngamma %.17g\n" % (init, tmp_path / "out", ic["h"], ic["gamma"]))
    r = subprocess.run([exe, "-c", str(cfg), "--ranks", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert r.returncode != 0, r.stdout[-2000:]
    assert "Slab decomposition over 2 GPUs" in r.stdout
    assert "mlh_create failed" in r.stdout
    assert os.listdir(tmp_path / "out") == []


@pytest.mark.parametrize("case,nranks", [("kh2d", 2), ("kh2d", 3), ("sedov3d", 2), ("sedov3d", 4)])
def test_cpp_launcher_shards_like_the_python_launcher(tmp_path, case, nranks):
    """Particles::slabParticles (C++ multi-rank launcher) == multigpu.shard (bench / test launcher): same cell-layer
    rule (Particles.cpp:279-302 on the slowest axis + mlh_slab_range), a partition of the particles."""
    from meshlesshydro_b200 import h5lite, ic as IC, multigpu
    ic = IC.kelvin_helmholtz(48, lattice=False) if case == "kh2d" else IC.sedov(24)
    init = str(tmp_path / "init.h5")
    h5lite.write_initial_conditions(init, ic)
    args = ["shard", init, "%.17g" % ic["h"], str(nranks)] + (["%.17g" % v for v in ic["box"]] if ic.get("box") is not None else [])
    r = _selftest(case, *args)
    seen = []
    for rank in range(nranks):
        ids_cpp = np.array([int(v) for v in r["shard%d" % rank].split()], dtype=np.int64)
        _, ids_py = multigpu.shard(ic, rank, nranks)
        assert np.array_equal(ids_cpp, ids_py.astype(np.int64)), rank
        seen.append(ids_cpp)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(len(ic["x"])))
