"""ctypes binding of the C ABI in include/mlh_gpu.h (the drop-in boundary).

This is the binding a maintainer of the reference would write; nothing here computes.  The shared
library is built in-tree (meshlesshydro_b200/libmlh_gpu.so, see csrc/Makefile / __graft_entry__.build)
and MUST be present: there is no CPU or PyTorch fallback -- a missing library or a missing CUDA
device raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MLH_GPU_LIB") or os.path.join(_HERE, "libmlh_gpu.so")  # override: A/B builds of the kernels

MLH_OK = 0
ABS_INT_TRUNC, ABS_FABS = 0, 1
Q13_ZERO_Z, Q13_GEOMETRIC = 0, 1
Q3_REFERENCE, Q3_FIXED = 0, 1
F_MAX_INTERACTIONS, F_OUT_OF_GRID, F_NEG_GHOST_PRESSURE, F_HALO_OVERFLOW, F_MIGRATION, F_VACUUM = 1, 2, 4, 8, 16, 32

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class MlhConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "dim", "periodic", "max_interactions", "slope_limiting", "pairwise_limiter", "meshless_finite_mass",
        "move_particles", "abs_mode", "q13_mode", "q3_mode", "symmetric_seam", "debug_capture", "first_order_quad_point",
        "reserved0")] + \
        [(n, C.c_double) for n in ("cfl", "beta", "psi1", "psi2", "kernel_size", "gamma")] + \
        [("box", C.c_double * 6), ("device", C.c_int), ("rank", C.c_int), ("nranks", C.c_int), ("capacity", C.c_long),
         ("stage_bytes", C.c_long)]


class MlhError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libmlh_gpu.so; raise (never fall back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MlhError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(the MFV path has no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sig = {
        "mlh_abi_version": (C.c_int, []),
        "mlh_default_config": (None, [C.POINTER(MlhConfig)]),
        "mlh_create": (C.c_int, [C.POINTER(MlhConfig), C.POINTER(vp)]),
        "mlh_destroy": (C.c_int, [vp]),
        "mlh_last_error": (C.c_char_p, [vp]),
        "mlh_error_flags": (C.c_uint, [vp]),
        "mlh_upload": (C.c_int, [vp, C.c_long] + [c_dp] * 8 + [c_ip]),
        "mlh_build_grid": (C.c_int, [vp]),
        "mlh_neighbours": (C.c_int, [vp]),
        "mlh_density_matrix": (C.c_int, [vp]),
        "mlh_gradients_limit": (C.c_int, [vp]),
        "mlh_timestep": (C.c_int, [vp, c_dp]),
        "mlh_flux_update": (C.c_int, [vp, C.c_double]),
        "mlh_prepare": (C.c_int, [vp, c_dp]),
        "mlh_advance": (C.c_int, [vp, C.c_double]),
        "mlh_step": (C.c_int, [vp, C.c_double, C.c_double, c_dp]),
        "mlh_download_state": (C.c_int, [vp] + [c_dp] * 8 + [c_ip]),
        "mlh_download_diag": (C.c_int, [vp, c_dp, c_dp, c_dp, c_ip]),
        "mlh_sums": (C.c_int, [vp, c_dp]),
        "mlh_num_particles": (C.c_long, [vp]),
        "mlh_grid_info": (C.c_int, [vp, c_ip, c_dp, c_dp]),
        "mlh_debug_fetch": (C.c_long, [vp, C.c_char_p, vp, C.c_long]),
        "mlh_stream": (vp, [vp]),
        "mlh_synchronize": (C.c_int, [vp]),
        "mlh_profile_enable": (C.c_int, [vp, C.c_int]),
        "mlh_profile_read": (C.c_int, [vp, C.c_int, C.POINTER(C.c_char_p), c_dp, C.POINTER(C.c_long)]),
        "mlh_launch_count": (C.c_long, [vp]),
        "mlh_timer_start": (C.c_int, [vp]),
        "mlh_timer_stop": (C.c_int, [vp, c_dp]),
        "mlh_comm_unique_id": (C.c_int, [C.c_char_p]),
        "mlh_comm_init": (C.c_int, [vp, C.c_char_p]),
        "mlh_slab_range": (C.c_int, [C.c_int, C.c_int, C.c_int, c_ip, c_ip]),
        "mlh_host_alloc": (C.c_int, [C.c_ulong, C.POINTER(vp)]),
        "mlh_host_free": (C.c_int, [vp]),
        "mlh_measure_fp64_peak": (C.c_int, [C.c_int, c_dp]),
        "mlh_riemann_faces": (C.c_int, [vp, C.c_long] + [c_dp] * 5),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "mlh_abi_version", "mlh_default_config", "mlh_create", "mlh_destroy", "mlh_last_error", "mlh_error_flags",
    "mlh_upload", "mlh_build_grid", "mlh_neighbours", "mlh_density_matrix", "mlh_gradients_limit", "mlh_timestep",
    "mlh_flux_update", "mlh_prepare", "mlh_advance", "mlh_step", "mlh_download_state", "mlh_download_diag", "mlh_sums",
    "mlh_num_particles", "mlh_grid_info", "mlh_debug_fetch", "mlh_stream", "mlh_synchronize", "mlh_profile_enable",
    "mlh_profile_read", "mlh_launch_count", "mlh_timer_start", "mlh_timer_stop", "mlh_comm_unique_id", "mlh_comm_init",
    "mlh_slab_range", "mlh_host_alloc", "mlh_host_free", "mlh_measure_fp64_peak", "mlh_riemann_faces",
]

_INT_FIELDS = {"cell", "noi", "noiGhosts", "sorted_index", "nnl", "nnlGhosts", "nnlGhostCodes", "num_faces", "face_pairs", "flux_symmetry"}


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def default_config():
    cfg = MlhConfig()
    load_library().mlh_default_config(C.byref(cfg))
    return cfg


# parameter.h presets of the reference's test cases
PRESETS = {
    # testcases/kelvin-helmholtz/parameter_long_run.h
    "kh2d": dict(dim=2, periodic=1, slope_limiting=1, pairwise_limiter=0, cfl=0.4, beta=4.0, psi1=0.5, psi2=0.25),
    # same, PERIODIC_BOUNDARIES 0 (fluid-block 2D)
    "fb2d": dict(dim=2, periodic=0, slope_limiting=1, pairwise_limiter=0, cfl=0.4, beta=4.0, psi1=0.5, psi2=0.25),
    # testcases/sedov/parameter.h
    "sedov3d": dict(dim=3, periodic=0, slope_limiting=1, pairwise_limiter=1, cfl=0.25, beta=1.0, psi1=0.5, psi2=0.25),
}


def make_config(preset, h, gamma, box=None, **over):
    cfg = default_config()
    vals = dict(PRESETS[preset])
    vals.update(over)
    for k, v in vals.items():
        setattr(cfg, k, v)
    cfg.kernel_size = h
    cfg.gamma = gamma
    if box is not None:
        for k, v in enumerate(box):
            cfg.box[k] = float(v)
    return cfg


def pinned_empty(n, dtype=np.float64):
    """numpy array over page-locked host memory from mlh_host_alloc (freed when the array is collected)."""
    lib = load_library()
    dt = np.dtype(dtype)
    nbytes = max(int(n), 1) * dt.itemsize
    ptr = C.c_void_p()
    rc = lib.mlh_host_alloc(nbytes, C.byref(ptr))
    if rc != MLH_OK:
        raise MlhError("mlh_host_alloc(%d) failed (%d): %s" % (nbytes, rc, lib.mlh_last_error(None).decode()))
    buf = (C.c_char * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(n))
    import weakref
    weakref.finalize(buf, lib.mlh_host_free, ptr)
    return arr


def fp64_peak_tflops(device=0):
    t = C.c_double()
    rc = load_library().mlh_measure_fp64_peak(device, C.byref(t))
    if rc != MLH_OK:
        raise MlhError("mlh_measure_fp64_peak failed (%d)" % rc)
    return t.value


class MfvGpu:
    """One GPU context of the MFV path (thin object wrapper over mlh_ctx)."""

    def __init__(self, cfg):
        self.lib = load_library()
        self.cfg = cfg
        self.ctx = C.c_void_p()
        rc = self.lib.mlh_create(C.byref(cfg), C.byref(self.ctx))
        if rc != MLH_OK:
            raise MlhError("mlh_create failed (%d): %s" % (rc, self.lib.mlh_last_error(None).decode()))
        self.D = cfg.dim
        self.N = 0

    def _check(self, rc, what):
        if rc < 0:
            raise MlhError("%s failed (%d): %s" % (what, rc, self.lib.mlh_last_error(self.ctx).decode()))
        return rc

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.lib.mlh_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ----
    def upload(self, ic, ids=None):
        arrs = [None if ic.get(k) is None else np.ascontiguousarray(ic[k], dtype=np.float64)
                for k in ("x", "y", "z", "vx", "vy", "vz", "m", "u")]
        self.N = len(arrs[0])
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
            idp = ids.ctypes.data_as(c_ip)
        self._check(self.lib.mlh_upload(self.ctx, self.N, *[_dp(a) for a in arrs], idp), "mlh_upload")

    def download_state(self, out=None):
        """current state in original particle order; `out` may hold preallocated (pinned) arrays"""
        n = self.lib.mlh_num_particles(self.ctx)
        names = ["x", "y", "z", "vx", "vy", "vz", "m", "u"]
        if out is None:
            out = {k: (np.empty(n) if (self.D == 3 or k not in ("z", "vz")) else None) for k in names}
        ids = out.get("ids")
        if ids is None:
            ids = np.empty(n, dtype=np.int32)
        self._check(self.lib.mlh_download_state(self.ctx, *[_dp(out[k]) for k in names], ids.ctypes.data_as(c_ip)),
                    "mlh_download_state")
        out["ids"] = ids
        return out

    def download_diag(self):
        n = self.lib.mlh_num_particles(self.ctx)
        rho, P, rg, noi = np.empty(n), np.empty(n), np.empty(n * self.D), np.empty(n, dtype=np.int32)
        self._check(self.lib.mlh_download_diag(self.ctx, _dp(rho), _dp(P), _dp(rg), noi.ctypes.data_as(c_ip)),
                    "mlh_download_diag")
        return dict(rho=rho, P=P, rhoGrad=rg.reshape(n, self.D), noi=noi)

    # ---- phases (MeshlessScheme::run order) ----
    def build_grid(self):
        self._check(self.lib.mlh_build_grid(self.ctx), "mlh_build_grid")

    def neighbours(self):
        self._check(self.lib.mlh_neighbours(self.ctx), "mlh_neighbours")

    def density_matrix(self):
        self._check(self.lib.mlh_density_matrix(self.ctx), "mlh_density_matrix")

    def gradients_limit(self):
        self._check(self.lib.mlh_gradients_limit(self.ctx), "mlh_gradients_limit")

    def timestep(self):
        dt = C.c_double()
        self._check(self.lib.mlh_timestep(self.ctx, C.byref(dt)), "mlh_timestep")
        return dt.value

    def flux_update(self, dt):
        self._check(self.lib.mlh_flux_update(self.ctx, dt), "mlh_flux_update")

    def prepare(self):
        dt = C.c_double()
        self._check(self.lib.mlh_prepare(self.ctx, C.byref(dt)), "mlh_prepare")
        return dt.value

    def advance(self, dt):
        self._check(self.lib.mlh_advance(self.ctx, dt), "mlh_advance")

    def step(self, dt_fixed=-1.0, dt_max=-1.0, want_dt=True):
        dt = C.c_double()
        self._check(self.lib.mlh_step(self.ctx, dt_fixed, dt_max, C.byref(dt) if want_dt else None), "mlh_step")
        return dt.value if want_dt else None

    # ---- results / diagnostics ----
    def sums(self):
        out = np.zeros(6)
        self._check(self.lib.mlh_sums(self.ctx, _dp(out)), "mlh_sums")
        return out

    def error_flags(self):
        return int(self.lib.mlh_error_flags(self.ctx))

    def grid(self):
        cells = (C.c_int * 3)()
        cs, b = np.zeros(3), np.zeros(6)
        self.lib.mlh_grid_info(self.ctx, cells, _dp(cs), _dp(b))
        return np.array(list(cells)), cs, b

    def fetch(self, name):
        n = self._check(self.lib.mlh_debug_fetch(self.ctx, name.encode(), None, 0), "mlh_debug_fetch(%s)" % name)
        if name in ("counters", "one_sided_pairs"):
            out = np.zeros(n, dtype=np.uint32)
        else:
            out = np.empty(n, dtype=np.int32 if name in _INT_FIELDS else np.float64)
        self._check(self.lib.mlh_debug_fetch(self.ctx, name.encode(), out.ctypes.data_as(C.c_void_p), n),
                    "mlh_debug_fetch(%s)" % name)
        return out

    # ---- measurement ----
    def riemann_faces(self, WR, WL, vFrame, Aij):
        """Riemann{WR, WL, vFrame, Aij}.exact for n faces (mlh_riemann_faces); arrays n x (D+2) / n x D."""
        WR = np.ascontiguousarray(WR, dtype=np.float64)
        WL = np.ascontiguousarray(WL, dtype=np.float64)
        vF = np.ascontiguousarray(vFrame, dtype=np.float64)
        A = np.ascontiguousarray(Aij, dtype=np.float64)
        n = WR.shape[0]
        F = np.empty_like(WR)
        self._check(self.lib.mlh_riemann_faces(self.ctx, n, _dp(WR), _dp(WL), _dp(vF), _dp(A), _dp(F)), "mlh_riemann_faces")
        return F

    def synchronize(self):
        self._check(self.lib.mlh_synchronize(self.ctx), "mlh_synchronize")

    def stream_handle(self):
        return self.lib.mlh_stream(self.ctx)

    def profile(self, on):
        self._check(self.lib.mlh_profile_enable(self.ctx, 1 if on else 0), "mlh_profile_enable")

    def profile_read(self):
        names = (C.c_char_p * 32)()
        ms = (C.c_double * 32)()
        ln = (C.c_long * 32)()
        n = self._check(self.lib.mlh_profile_read(self.ctx, 32, names, ms, ln), "mlh_profile_read")
        return {names[k].decode(): (ms[k], ln[k]) for k in range(n)}

    def launch_count(self):
        return int(self.lib.mlh_launch_count(self.ctx))

    def timer_start(self):
        self._check(self.lib.mlh_timer_start(self.ctx), "mlh_timer_start")

    def timer_stop(self):
        ms = C.c_double()
        self._check(self.lib.mlh_timer_stop(self.ctx, C.byref(ms)), "mlh_timer_stop")
        return ms.value
