"""ctypes loaders for the two CPU checkers (TEST INFRASTRUCTURE, never used by the product):

* ``Oracle``    -- oracle/liboracle.so, the in-repo C restatement (oracle/mfv_oracle.c)
* ``Reference`` -- oracle/_ref/libref_<variant>.so, the reference's own sources compiled
                   by oracle/ref_build/Makefile (prebuilt files travel to the GPU box)

Both expose create/step/fetch with identical semantics so tests can diff them.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
c_dp = C.POINTER(C.c_double)


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


INT_FIELDS = {"cell", "noi", "nnl", "noiGhosts", "nnlGhosts", "ghostMap", "ghost_parent", "ghost_N",
              "one_sided", "err_flags"}


class OrcConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("dim", "periodic", "max_ni", "max_gi", "slope_limiting", "pairwise", "mfm",
                                       "move_particles", "abs_mode", "q13_mode", "q3_mode", "quad_point_h4")] + \
               [(n, C.c_double) for n in ("cfl", "beta", "psi1", "psi2", "h", "gamma")] + [("box", C.c_double * 6)]


# parameter.h presets of the reference's test cases (values: see oracle/ref_build/Makefile variants)
PRESETS = {
    # testcases/kelvin-helmholtz/parameter_long_run.h
    "kh2d": dict(dim=2, periodic=1, max_ni=400, max_gi=300, slope_limiting=1, pairwise=0, cfl=0.4, beta=4.0,
                 psi1=0.5, psi2=0.25),
    # same with PERIODIC_BOUNDARIES 0 (fluid-block 2D)
    "fb2d": dict(dim=2, periodic=0, max_ni=400, max_gi=300, slope_limiting=1, pairwise=0, cfl=0.4, beta=4.0,
                 psi1=0.5, psi2=0.25),
    # testcases/sedov/parameter.h
    "sedov3d": dict(dim=3, periodic=0, max_ni=400, max_gi=300, slope_limiting=1, pairwise=1, cfl=0.25, beta=1.0,
                    psi1=0.5, psi2=0.25),
}


def make_config(preset, h, gamma, box=None, abs_mode=0, q13_mode=0, q3_mode=0, **over):
    p = dict(PRESETS[preset])
    p.update(over)
    cfg = OrcConfig()
    cfg.dim, cfg.periodic = p["dim"], p["periodic"]
    cfg.max_ni, cfg.max_gi = p["max_ni"], p["max_gi"]
    cfg.slope_limiting, cfg.pairwise = p["slope_limiting"], p["pairwise"]
    cfg.mfm, cfg.move_particles = p.get("mfm", 0), p.get("move_particles", 1)
    cfg.abs_mode, cfg.q13_mode, cfg.q3_mode = abs_mode, q13_mode, q3_mode
    cfg.quad_point_h4 = p.get("quad_point_h4", 0)
    cfg.cfl, cfg.beta, cfg.psi1, cfg.psi2 = p["cfl"], p["beta"], p["psi1"], p["psi2"]
    cfg.h, cfg.gamma = h, gamma
    for k in range(6):
        cfg.box[k] = 0.0
    if box is not None:
        for k, v in enumerate(box):
            cfg.box[k] = float(v)
    return cfg


class _Base:
    prefix = ""

    def _bind(self, lib):
        p = self.prefix
        self._step = getattr(lib, p + "step")
        self._step.restype = C.c_double
        self._step.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
        self._fetch = getattr(lib, p + "fetch")
        self._fetch.restype = C.c_long
        self._fetch.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        self._sums = getattr(lib, p + "sums")
        self._sums.argtypes = [C.c_void_p, c_dp]
        self._grid = getattr(lib, p + "grid")
        self._grid.argtypes = [C.c_void_p, C.POINTER(C.c_int), c_dp, c_dp]
        self._dtcfl = getattr(lib, p + "last_dt_cfl")
        self._dtcfl.restype = C.c_double
        self._dtcfl.argtypes = [C.c_void_p]
        self._phase = getattr(lib, p + "phase_seconds")
        self._phase.argtypes = [C.c_void_p, c_dp]
        self._destroy = getattr(lib, p + "destroy")
        self._destroy.argtypes = [C.c_void_p]

    def step(self, dt_fixed=-1.0, dt_max=-1.0, stop_after=0):
        return self._step(self.ctx, dt_fixed, dt_max, stop_after)

    @property
    def max_ni(self):  # slots per particle of the per-slot arrays (MAX_NUM_INTERACTIONS)
        return self.cfg.max_ni if hasattr(self, "cfg") else self.info["max_ni"]

    @property
    def max_gi(self):  # MAX_NUM_GHOST_INTERACTIONS
        return self.cfg.max_gi if hasattr(self, "cfg") else self.info["max_gi"]

    def fetch(self, name):
        n = self._fetch(self.ctx, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.int32 if name in INT_FIELDS else np.float64)
        self._fetch(self.ctx, name.encode(), out.ctypes.data_as(C.c_void_p))
        return out

    def sums(self):
        out = np.zeros(6)
        self._sums(self.ctx, _dp(out))
        return out

    def grid(self):
        cells = (C.c_int * 3)()
        cs = np.zeros(3)
        b = np.zeros(6)
        self._grid(self.ctx, cells, _dp(cs), _dp(b))
        return np.array(list(cells)), cs, b

    def dt_cfl(self):
        return self._dtcfl(self.ctx)

    def phase_seconds(self):
        out = np.zeros(11)
        self._phase(self.ctx, _dp(out))
        return out

    def close(self):
        if getattr(self, "ctx", None):
            self._destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Oracle(_Base):
    prefix = "orc_"
    path = os.path.join(ROOT, "oracle", "liboracle.so")

    def __init__(self, cfg, ic):
        self.lib = C.CDLL(self.path)
        self._bind(self.lib)
        self.cfg = cfg
        self.N = len(ic["x"])
        self.lib.orc_create.restype = C.c_void_p
        self.lib.orc_create.argtypes = [C.POINTER(OrcConfig), C.c_int] + [c_dp] * 8
        arrs = [np.ascontiguousarray(ic.get(k), dtype=np.float64) if ic.get(k) is not None else None
                for k in ("x", "y", "z", "vx", "vy", "vz", "m", "u")]
        self.ctx = self.lib.orc_create(C.byref(cfg), self.N, *[_dp(a) for a in arrs])


class Reference(_Base):
    prefix = "ref_"

    @staticmethod
    def lib_path(variant):
        return os.path.join(ROOT, "oracle", "_ref", "libref_%s.so" % variant)

    @classmethod
    def available(cls, variant):
        return os.path.exists(cls.lib_path(variant))

    def __init__(self, variant, ic):
        self.lib = C.CDLL(self.lib_path(variant))
        self._bind(self.lib)
        iv = (C.c_int * 12)()
        dv = (C.c_double * 4)()
        self.lib.ref_info(iv, dv)
        self.info = dict(dim=iv[0], periodic=iv[1], max_ni=iv[2], max_gi=iv[3], pairwise=iv[4], slope_limiting=iv[5],
                         mfm=iv[6], flux_sym=iv[7], move=iv[8], first_order=iv[9], adaptive=iv[10], fabs=iv[11],
                         cfl=dv[0], beta=dv[1], psi1=dv[2], psi2=dv[3])
        assert self.info["dim"] == ic["dim"] and self.info["periodic"] == ic["periodic"], (self.info, ic["dim"])
        self.N = len(ic["x"])
        self.lib.ref_create.restype = C.c_void_p
        self.lib.ref_create.argtypes = [C.c_int] + [c_dp] * 9 + [C.c_double, C.c_double]
        arrs = [np.ascontiguousarray(ic.get(k), dtype=np.float64) if ic.get(k) is not None else None
                for k in ("x", "y", "z", "vx", "vy", "vz", "m", "u")]
        box = None if ic.get("box") is None else np.ascontiguousarray(ic["box"], dtype=np.float64)
        self.ctx = self.lib.ref_create(self.N, *[_dp(a) for a in arrs], _dp(box), ic["h"], ic["gamma"])

    def time_steps(self, nsteps):
        self.lib.ref_time_steps.restype = C.c_double
        self.lib.ref_time_steps.argtypes = [C.c_void_p, C.c_int, c_dp]
        per = np.zeros(nsteps)
        med = self.lib.ref_time_steps(self.ctx, nsteps, _dp(per))
        return med, per
