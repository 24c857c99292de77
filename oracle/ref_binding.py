"""ctypes binding of oracle/_ref/libgtop_reference.so (TEST INFRASTRUCTURE ONLY).

That library is the REFERENCE'S OWN hot-path sources (src/sdf_map.cpp, src/qp_generator.cpp,
src/grad_traj_optimizer.cpp), compiled unmodified against the stand-in headers of oracle/shim/ by
`make -C oracle reference` (only where /root/reference exists), plus the C wrapper oracle/ref_capi.cpp.
It pins the oracle restatement to the reference itself: tests/test_reference_build.py compares the two
directly, and tests/golden/make_reference_golden.py stores its outputs for the GPU box, which has no
/root/reference.  Nothing in the product may load it.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PARAM_ORDER = ("alg", "time_limit_1", "time_limit_2", "dt", "ws", "wc", "alpha", "r", "d0", "alpha_v", "r_v", "v0",
               "alpha_a", "r_a", "a0", "bos", "vos", "aos", "mean_v", "mean_a", "init_time")
_ALIASES = {"w_smooth": "ws", "w_collision": "wc", "algorithm": "alg", "deltat": "dt"}
_DEFAULTS = dict(alg=24, time_limit_1=0.04, time_limit_2=0.06, dt=0.2, ws=1.0, wc=1.0, alpha=1.0, r=1.0, d0=1.0,
                 alpha_v=0.0, r_v=1.0, v0=1.0, alpha_a=0.0, r_a=1.0, a0=1.0, bos=3.0, vos=8.0, aos=10.0, mean_v=1.0,
                 mean_a=1.0, init_time=0.0)


def lib_path():
    return os.path.join(_HERE, "_ref", "libgtop_reference.so")


def available():
    return os.path.exists(lib_path())


_lib = None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError("oracle/_ref/libgtop_reference.so is not built (needs /root/reference: make -C oracle reference)")
    # the nlopt.hpp stand-in dlopens the NLopt 2.5.0 binary the reference vendors
    os.environ.setdefault("GTOP_NLOPT_LIB", os.path.join(_HERE, "_ref", "libnlopt.so.0"))
    L = C.CDLL(lib_path())
    vp, dp, i64, ip = C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int)
    sig = {
        "ref_last_error": (C.c_char_p, []),
        "ref_map_create": (vp, [dp, C.c_double, dp]),
        "ref_map_destroy": (None, [vp]),
        "ref_map_grid": (None, [vp, ip]),
        "ref_map_reset": (None, [vp]),
        "ref_map_reset_range": (None, [vp, dp, dp]),
        "ref_map_set_update_range": (None, [vp, dp, dp]),
        "ref_map_update_range": (None, [vp, ip]),
        "ref_map_set_occupancy": (None, [vp, dp, i64]),
        "ref_map_update_esdf": (None, [vp]),
        "ref_map_get_distance": (None, [vp, dp]),
        "ref_map_get_occupancy": (None, [vp, C.POINTER(C.c_int32)]),
        "ref_map_query": (None, [vp, dp, i64, dp, dp]),
        "ref_map_max_distance": (C.c_double, [vp]),
        "ref_opt_create": (vp, [dp]),
        "ref_opt_destroy": (None, [vp]),
        "ref_opt_init_map": (C.c_int, [vp, dp, dp, C.c_double]),
        "ref_opt_update_map": (C.c_int, [vp, dp, i64]),
        "ref_opt_get_distance": (None, [vp, dp]),
        "ref_opt_set_distance": (None, [vp, dp]),
        "ref_opt_set_path": (C.c_int, [vp, dp, C.c_int]),
        "ref_opt_set_kino_path": (C.c_int, [vp, dp, dp, dp, dp, C.c_int]),
        "ref_opt_dims": (None, [vp, ip]),
        "ref_opt_seg_time": (None, [vp, dp]),
        "ref_opt_L": (None, [vp, dp]),
        "ref_opt_R": (None, [vp, dp]),
        "ref_opt_Df": (None, [vp, dp]),
        "ref_opt_Dp": (None, [vp, dp]),
        "ref_opt_coeff": (None, [vp, dp]),
        "ref_opt_x0": (None, [vp, dp]),
        "ref_opt_eval": (C.c_int, [vp, C.c_int, dp, i64, dp, dp]),
        "ref_opt_coeff_from_x": (None, [vp, dp, dp]),
        "ref_opt_optimize": (C.c_int, [vp, C.c_int, C.c_int, dp, ip]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


class ReferenceMap:
    """The reference's SDFMap (include/grad_traj_optimization/sdf_map.h), compiled from its own source."""

    def __init__(self, origin, resolution, size):
        o, s = np.asarray(origin, dtype=np.float64), np.asarray(size, dtype=np.float64)
        self.h = lib().ref_map_create(_dp(o), float(resolution), _dp(s))
        if not self.h:
            raise RuntimeError(lib().ref_last_error().decode())
        g = (C.c_int * 3)()
        lib().ref_map_grid(self.h, g)
        self.grid = tuple(int(v) for v in g)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_map_destroy(self.h)
            self.h = None

    @property
    def nvox(self):
        return self.grid[0] * self.grid[1] * self.grid[2]

    def reset(self):
        lib().ref_map_reset(self.h)

    def reset_range(self, lo, hi):
        a, b = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
        lib().ref_map_reset_range(self.h, _dp(a), _dp(b))

    def set_update_range(self, lo, hi):
        a, b = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
        lib().ref_map_set_update_range(self.h, _dp(a), _dp(b))

    def update_range(self):
        out = (C.c_int * 6)()
        lib().ref_map_update_range(self.h, out)
        v = np.array(list(out), dtype=np.int32)
        return v[:3], v[3:]

    def set_occupancy(self, pts):
        p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        lib().ref_map_set_occupancy(self.h, _dp(p), p.shape[0])

    def update_esdf(self):
        lib().ref_map_update_esdf(self.h)

    def distance(self):
        out = np.empty(self.nvox)
        lib().ref_map_get_distance(self.h, _dp(out))
        return out.reshape(self.grid)

    def occupancy(self):
        out = np.empty(self.nvox, dtype=np.int32)
        lib().ref_map_get_occupancy(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out.reshape(self.grid)

    def query(self, pos):
        p = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        dist, grad = np.empty(p.shape[0]), np.empty((p.shape[0], 3))
        lib().ref_map_query(self.h, _dp(p), p.shape[0], _dp(dist), _dp(grad))
        return dist, grad

    def max_distance(self):
        return float(lib().ref_map_max_distance(self.h))


class ReferenceOptimizer:
    """The reference's GradTrajOptimizer, compiled from its own source (parameters through the ros::param stand-in)."""

    def __init__(self, params):
        d = dict(_DEFAULTS)
        for k, v in params.items():
            k = _ALIASES.get(k, k)
            if k in d:
                d[k] = float(v)
        self.params = d
        vec = np.array([d[k] for k in PARAM_ORDER], dtype=np.float64)
        self.h = lib().ref_opt_create(_dp(vec))
        if not self.h:
            raise RuntimeError(lib().ref_last_error().decode())
        self.grid = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_opt_destroy(self.h)
            self.h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(lib().ref_last_error().decode())

    def init_map(self, size, origin, resolution):
        s, o = np.asarray(size, dtype=np.float64), np.asarray(origin, dtype=np.float64)
        self._check(lib().ref_opt_init_map(self.h, _dp(s), _dp(o), float(resolution)))
        self.grid = tuple(int(np.ceil(s[i] / resolution)) for i in range(3))

    def update_map(self, obstacles):
        p = np.ascontiguousarray(obstacles, dtype=np.float64).reshape(-1, 3)
        self._check(lib().ref_opt_update_map(self.h, _dp(p), p.shape[0]))

    def distance(self):
        out = np.empty(int(np.prod(self.grid)))
        lib().ref_opt_get_distance(self.h, _dp(out))
        return out.reshape(self.grid)

    def set_distance(self, d):
        d = np.ascontiguousarray(d, dtype=np.float64).reshape(-1)
        assert d.size == int(np.prod(self.grid))
        lib().ref_opt_set_distance(self.h, _dp(d))

    def set_path(self, waypoints):
        w = np.ascontiguousarray(waypoints, dtype=np.float64).reshape(-1, 3)
        self._check(lib().ref_opt_set_path(self.h, _dp(w), w.shape[0]))
        self._dims()

    def set_kino_path(self, pos, vel, acc, time):
        p, v, a = (np.ascontiguousarray(t, dtype=np.float64).reshape(-1, 3) for t in (pos, vel, acc))
        t = np.ascontiguousarray(time, dtype=np.float64)
        self._check(lib().ref_opt_set_kino_path(self.h, _dp(p), _dp(v), _dp(a), _dp(t), t.size))
        self._dims()

    def _dims(self):
        out = (C.c_int * 3)()
        lib().ref_opt_dims(self.h, out)
        self.m, self.num_dp, self.num_df = int(out[0]), int(out[1]), int(out[2])
        self.n = 3 * self.num_dp

    def _get(self, fn, shape, order="F"):
        out = np.empty(int(np.prod(shape)))
        getattr(lib(), fn)(self.h, _dp(out))
        return out.reshape(shape, order=order)

    def seg_time(self):
        return self._get("ref_opt_seg_time", (self.m,))

    def L(self):
        return self._get("ref_opt_L", (6 * self.m, 3 * self.m + 3))

    def R(self):
        return self._get("ref_opt_R", (3 * self.m + 3, 3 * self.m + 3))

    def Df(self):
        return self._get("ref_opt_Df", (3, 6))

    def coeff(self):
        return self._get("ref_opt_coeff", (self.m, 18))

    def x0(self):
        return self._get("ref_opt_x0", (self.n,))

    def coeff_from_x(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(self.m * 18)
        lib().ref_opt_coeff_from_x(self.h, _dp(x), _dp(out))
        return out.reshape((self.m, 18), order="F")

    def eval_batch(self, X, step=2):
        """costFunc (grad_traj_optimizer.cpp:554-562) on every row of X."""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.n)
        cost, grad = np.empty(X.shape[0]), np.empty_like(X)
        self._check(lib().ref_opt_eval(self.h, int(step), _dp(X), X.shape[0], _dp(cost), _dp(grad)))
        return cost, grad

    def optimize(self, step, maxeval):
        """optimizeTrajectory(step) (:128-243) stopped after `maxeval` objective evaluations; returns (x, evaluations)."""
        x = np.empty(self.n)
        ev = C.c_int(0)
        self._check(lib().ref_opt_optimize(self.h, int(step), int(maxeval), _dp(x), C.byref(ev)))
        return x, int(ev.value)
