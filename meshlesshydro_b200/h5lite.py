"""ctypes binding of H5Lite (meshlesshydro_b200/host/src/H5Lite.cpp): the self-contained HDF5 subset reader/writer
the host layer uses for initial conditions and snapshots (h5py / libhdf5 are not available in this image).

    write_file(path, {"x": array, ...})   datasets in the root group; float64, int32 or int8 arrays
    read_file(path) -> {name: array}      float datasets as float64, integer datasets as int32
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "host", "libh5lite.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s not built: run `make -C meshlesshydro_b200/host` (or __graft_entry__.build())" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.h5lite_last_error.restype = C.c_char_p
        L.h5lite_open.restype = vp
        L.h5lite_open.argtypes = [C.c_char_p]
        L.h5lite_close.argtypes = [vp]
        L.h5lite_num_datasets.argtypes = [vp]
        L.h5lite_dataset_name.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
        L.h5lite_info.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.h5lite_read_f64.restype = C.c_longlong
        L.h5lite_read_f64.argtypes = [vp, C.c_char_p, vp, C.c_longlong]
        L.h5lite_read_i32.restype = C.c_longlong
        L.h5lite_read_i32.argtypes = [vp, C.c_char_p, vp, C.c_longlong]
        L.h5lite_create.restype = vp
        L.h5lite_create.argtypes = [C.c_char_p]
        L.h5lite_write.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_ulonglong), C.c_int, vp]
        L.h5lite_finish.argtypes = [vp]
        _lib = L
    return _lib


class H5Error(RuntimeError):
    pass


def write_file(path, datasets):
    L = _load()
    w = L.h5lite_create(os.fsencode(path))
    for name, arr in datasets.items():
        a = np.asarray(arr)
        if a.dtype.kind == "f":
            a, t = np.ascontiguousarray(a, dtype=np.float64), 0
        elif a.dtype == np.int8:
            a, t = np.ascontiguousarray(a), 2
        else:
            a, t = np.ascontiguousarray(a, dtype=np.int32), 1
        shape = a.shape if a.ndim else (1,)
        dims = (C.c_ulonglong * len(shape))(*shape)
        if L.h5lite_write(w, name.encode(), len(shape), dims, t, a.ctypes.data) != 0:
            err = L.h5lite_last_error().decode()
            L.h5lite_finish(w)
            raise H5Error(err)
    if L.h5lite_finish(w) != 0:
        raise H5Error(L.h5lite_last_error().decode())


def read_file(path):
    L = _load()
    f = L.h5lite_open(os.fsencode(path))
    if not f:
        raise H5Error(L.h5lite_last_error().decode())
    out = {}
    try:
        for k in range(L.h5lite_num_datasets(f)):
            buf = C.create_string_buffer(512)
            L.h5lite_dataset_name(f, k, buf, 512)
            rank, kind, elem = C.c_int(), C.c_int(), C.c_int()
            dims = (C.c_ulonglong * 8)()
            if L.h5lite_info(f, buf.value, C.byref(rank), dims, C.byref(kind), C.byref(elem)) != 0:
                raise H5Error(L.h5lite_last_error().decode())
            shape = tuple(int(dims[i]) for i in range(rank.value))
            n = int(np.prod(shape)) if shape else 1
            if kind.value == 0:
                a = np.empty(n, dtype=np.float64)
                got = L.h5lite_read_f64(f, buf.value, a.ctypes.data, n)
            else:
                a = np.empty(n, dtype=np.int32)
                got = L.h5lite_read_i32(f, buf.value, a.ctypes.data, n)
            if got != n:
                raise H5Error("%s: %s" % (buf.value.decode(), L.h5lite_last_error().decode()))
            out[buf.value.decode()] = a.reshape(shape)
    finally:
        L.h5lite_close(f)
    return out


def write_initial_conditions(path, ic):
    """IC file with the datasets InitialDistribution reads (/m /x /v /u /materialId; generateIC.py:97-111)."""
    D = ic["dim"]
    cols = ["x", "y", "z"][:D]
    vcols = ["vx", "vy", "vz"][:D]
    write_file(path, {
        "m": ic["m"], "u": ic["u"],
        "x": np.stack([ic[c] for c in cols], axis=1),
        "v": np.stack([ic[c] for c in vcols], axis=1),
        "materialId": np.zeros(len(ic["m"]), dtype=np.int8),
    })
