"""ctypes binding of oracle/_ref/libgtop_oracle.so (TEST INFRASTRUCTURE ONLY).

Mirrors oracle/gtop_oracle_capi.cpp.  The library is built by oracle/Makefile (g++ -O2
-ffp-contract=off), which `build()` runs on demand when a compiler is available.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "_ref", "libgtop_oracle.so")


def nlopt_path():
    """Path of the vendored NLopt 2.5.0 binary copy (or None when it is not available)."""
    p = os.environ.get("GTOP_NLOPT_LIB") or os.path.join(_HERE, "_ref", "libnlopt.so.0")
    return p if os.path.exists(p) else None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("gtop_oracle.cpp", "gtop_oracle_capi.cpp", "gtop_oracle.hpp")]
    out = lib_path()
    stale = (not os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return out


class OracleParams(C.Structure):
    """Field order mirrors gtop_oracle::Params."""
    _fields_ = [(k, C.c_double) for k in (
        "time_limit_1", "time_limit_2", "deltat", "w_smooth", "w_collision", "alpha", "r", "d0",
        "alpha_v", "r_v", "v0", "alpha_a", "r_a", "a0", "bos", "vos", "aos", "mean_v", "mean_a", "init_time")] + [
        (k, C.c_int32) for k in ("algorithm", "samples_per_segment", "vel_acc_penalty", "exact_gradient")]

    @classmethod
    def from_dict(cls, d):
        p = cls(time_limit_1=0.04, time_limit_2=0.06, deltat=0.2, w_smooth=1.0, w_collision=1.0, alpha=1.0, r=1.0,
                d0=1.0, alpha_v=0.0, r_v=1.0, v0=1.0, alpha_a=0.0, r_a=1.0, a0=1.0, bos=3.0, vos=8.0, aos=10.0,
                mean_v=1.0, mean_a=1.0, init_time=0.0, algorithm=24, samples_per_segment=30, vel_acc_penalty=0,
                exact_gradient=0)
        names = {f[0] for f in cls._fields_}
        for k, v in d.items():
            if k in names:
                setattr(p, k, v)
        return p


_lib = None
_NLOPT_FUNC = C.CFUNCTYPE(C.c_double, C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(lib_path())
    vp, dp, i64 = C.c_void_p, C.POINTER(C.c_double), C.c_int64
    sig = {
        "gto_last_error": (C.c_char_p, []),
        "gto_map_create": (vp, [dp, C.c_double, dp]),
        "gto_map_destroy": (None, [vp]),
        "gto_map_grid": (None, [vp, C.POINTER(C.c_int)]),
        "gto_map_reset": (None, [vp]),
        "gto_map_reset_range": (None, [vp, dp, dp]),
        "gto_map_set_update_range": (None, [vp, dp, dp]),
        "gto_map_update_range": (None, [vp, C.POINTER(C.c_int32)]),
        "gto_map_set_occupancy": (None, [vp, dp, i64]),
        "gto_map_update_esdf": (None, [vp]),
        "gto_map_get_distance": (None, [vp, dp]),
        "gto_map_set_distance": (None, [vp, dp]),
        "gto_map_get_occupancy": (None, [vp, C.POINTER(C.c_int32)]),
        "gto_map_set_occupancy_buffer": (None, [vp, C.POINTER(C.c_int32)]),
        "gto_map_brute_edt": (None, [vp, dp]),
        "gto_map_query": (i64, [vp, dp, i64, dp, dp]),
        "gto_opt_create": (vp, [C.POINTER(OracleParams)]),
        "gto_opt_destroy": (None, [vp]),
        "gto_opt_set_map": (None, [vp, vp]),
        "gto_opt_set_params": (None, [vp, C.POINTER(OracleParams)]),
        "gto_opt_set_path": (C.c_int, [vp, dp, C.c_int]),
        "gto_opt_set_kino_path": (C.c_int, [vp, dp, dp, dp, dp, C.c_int]),
        "gto_opt_dims": (None, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "gto_opt_get_L": (None, [vp, dp]),
        "gto_opt_get_R": (None, [vp, dp]),
        "gto_opt_get_A": (None, [vp, dp]),
        "gto_opt_get_C": (None, [vp, dp]),
        "gto_opt_get_Df": (None, [vp, dp]),
        "gto_opt_get_seg_time": (None, [vp, dp]),
        "gto_opt_get_x0": (None, [vp, dp]),
        "gto_opt_get_bounds": (None, [vp, dp, dp]),
        "gto_opt_get_coeff": (None, [vp, dp]),
        "gto_opt_coeff_from_x": (None, [vp, dp, dp]),
        "gto_opt_eval": (C.c_double, [vp, dp, dp, C.c_int, C.POINTER(i64), dp]),
        "gto_opt_eval_batch": (C.c_double, [vp, dp, i64, dp, dp, C.c_int, C.c_int, C.POINTER(i64)]),
        "gto_opt_solve": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_char_p, C.c_void_p, vp, dp, dp,
                                    C.POINTER(C.c_int)]),
        "gto_opt_cost_curve": (i64, [vp, dp, i64]),
        "gto_traj_metrics": (i64, [dp, dp, C.c_int, dp, dp, i64]),
        "gto_edt_env_query": (None, [vp, dp, C.c_int, dp, i64, C.c_double, dp, dp, dp]),
        "gto_hardware_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


class OracleMap:
    """SDFMap restatement (sdf_map.cpp)."""

    def __init__(self, origin, resolution, size):
        o = np.asarray(origin, dtype=np.float64)
        s = np.asarray(size, dtype=np.float64)
        self.h = lib().gto_map_create(_dp(o), float(resolution), _dp(s))
        if not self.h:
            raise RuntimeError(lib().gto_last_error().decode())
        g = (C.c_int * 3)()
        lib().gto_map_grid(self.h, g)
        self.grid = tuple(int(v) for v in g)
        self.origin, self.resolution, self.size = o.copy(), float(resolution), s.copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().gto_map_destroy(self.h)
            self.h = None

    @property
    def nvox(self):
        return self.grid[0] * self.grid[1] * self.grid[2]

    def reset(self):
        lib().gto_map_reset(self.h)

    def reset_range(self, min_pos, max_pos):
        """SDFMap::resetBuffer(min_pos, max_pos) (sdf_map.cpp:28-53)."""
        a, b = np.asarray(min_pos, dtype=np.float64), np.asarray(max_pos, dtype=np.float64)
        lib().gto_map_reset_range(self.h, _dp(a), _dp(b))

    def set_update_range(self, min_pos, max_pos):
        """SDFMap::setUpdateRange (sdf_map.cpp:244-264): the window of the NEXT update_esdf only (:366-367)."""
        a, b = np.asarray(min_pos, dtype=np.float64), np.asarray(max_pos, dtype=np.float64)
        lib().gto_map_set_update_range(self.h, _dp(a), _dp(b))

    def update_range(self):
        out = np.zeros(6, dtype=np.int32)
        lib().gto_map_update_range(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out[:3].copy(), out[3:].copy()

    def set_occupancy(self, pts):
        p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        lib().gto_map_set_occupancy(self.h, _dp(p), p.shape[0])

    def update_esdf(self):
        lib().gto_map_update_esdf(self.h)

    def distance(self):
        out = np.empty(self.nvox, dtype=np.float64)
        lib().gto_map_get_distance(self.h, _dp(out))
        return out.reshape(self.grid)

    def set_distance(self, d):
        d = np.ascontiguousarray(d, dtype=np.float64).reshape(-1)
        assert d.size == self.nvox
        lib().gto_map_set_distance(self.h, _dp(d))

    def occupancy(self):
        out = np.empty(self.nvox, dtype=np.int32)
        lib().gto_map_get_occupancy(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out.reshape(self.grid)

    def set_occupancy_buffer(self, occ):
        o = np.ascontiguousarray(occ, dtype=np.int32).reshape(-1)
        assert o.size == self.nvox
        lib().gto_map_set_occupancy_buffer(self.h, o.ctypes.data_as(C.POINTER(C.c_int32)))

    def query_dynamic(self, pos, boxes, time):
        """EDTEnvironment::evaluateEDTWithGrad / evaluateCoarseEDT (edt_environment.cpp:75-137): boxes [nbox, 9] =
        (c0, c1, scale) of constant-velocity boxes.  Returns (dist[n], grad[n,3], coarse[n])."""
        p = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 9)
        dist, grad, coarse = np.empty(p.shape[0]), np.empty((p.shape[0], 3)), np.empty(p.shape[0])
        lib().gto_edt_env_query(self.h, _dp(b), b.shape[0], _dp(p), p.shape[0], float(time), _dp(dist), _dp(grad), _dp(coarse))
        return dist, grad, coarse

    def brute_edt(self):
        out = np.empty(self.nvox, dtype=np.float64)
        lib().gto_map_brute_edt(self.h, _dp(out))
        return out.reshape(self.grid)

    def query(self, pos):
        p = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        dist = np.empty(p.shape[0], dtype=np.float64)
        grad = np.empty((p.shape[0], 3), dtype=np.float64)
        oob = lib().gto_map_query(self.h, _dp(p), p.shape[0], _dp(dist), _dp(grad))
        return dist, grad, int(oob)


class OracleOptimizer:
    """GradTrajOptimizer restatement (grad_traj_optimizer.cpp + qp_generator.cpp)."""

    def __init__(self, params, omap=None):
        self.params = params if isinstance(params, OracleParams) else OracleParams.from_dict(params)
        self.h = lib().gto_opt_create(C.byref(self.params))
        self.map = None
        if omap is not None:
            self.set_map(omap)

    def __del__(self):
        if getattr(self, "h", None):
            lib().gto_opt_destroy(self.h)
            self.h = None

    def set_map(self, omap):
        self.map = omap  # keep alive
        lib().gto_opt_set_map(self.h, omap.h)

    def set_params(self, params):
        self.params = params if isinstance(params, OracleParams) else OracleParams.from_dict(params)
        lib().gto_opt_set_params(self.h, C.byref(self.params))

    def set_path(self, waypoints):
        w = np.ascontiguousarray(waypoints, dtype=np.float64).reshape(-1, 3)
        if lib().gto_opt_set_path(self.h, _dp(w), w.shape[0]) != 0:
            raise RuntimeError(lib().gto_last_error().decode())
        self._dims()

    def set_kino_path(self, pos, vel, acc, time):
        p, v, a = (np.ascontiguousarray(t, dtype=np.float64).reshape(-1, 3) for t in (pos, vel, acc))
        t = np.ascontiguousarray(time, dtype=np.float64)
        if lib().gto_opt_set_kino_path(self.h, _dp(p), _dp(v), _dp(a), _dp(t), t.size) != 0:
            raise RuntimeError(lib().gto_last_error().decode())
        self._dims()

    def _dims(self):
        m, p = C.c_int(), C.c_int()
        lib().gto_opt_dims(self.h, C.byref(m), C.byref(p))
        self.m, self.num_dp = m.value, p.value
        self.n = 3 * self.num_dp

    def _get(self, fn, shape, order="C"):
        out = np.empty(int(np.prod(shape)), dtype=np.float64)
        getattr(lib(), fn)(self.h, _dp(out))
        return out.reshape(shape, order=order)

    # matrices (column-major in the library, returned as 2-D numpy arrays with the right values)
    def L(self):
        return self._get("gto_opt_get_L", (6 * self.m, 3 * self.m + 3), order="F")

    def R(self):
        return self._get("gto_opt_get_R", (3 * self.m + 3, 3 * self.m + 3), order="F")

    def A(self):
        return self._get("gto_opt_get_A", (6 * self.m, 6 * self.m), order="F")

    def Cmat(self):
        return self._get("gto_opt_get_C", (3 * self.m + 3, 6 * self.m), order="F")

    def Df(self):
        return self._get("gto_opt_get_Df", (3, 6))

    def seg_time(self):
        return self._get("gto_opt_get_seg_time", (self.m,))

    def x0(self):
        return self._get("gto_opt_get_x0", (self.n,))

    def bounds(self):
        lb = np.empty(self.n)
        ub = np.empty(self.n)
        lib().gto_opt_get_bounds(self.h, _dp(lb), _dp(ub))
        return lb, ub

    def coeff(self):
        return self._get("gto_opt_get_coeff", (self.m, 18))

    def coeff_from_x(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty((self.m, 18))
        lib().gto_opt_coeff_from_x(self.h, _dp(x), _dp(out))
        return out

    def eval(self, x, step=2, want_terms=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.n
        g = np.empty(self.n)
        oob = C.c_int64(0)
        terms = np.zeros(4)
        c = lib().gto_opt_eval(self.h, _dp(x), _dp(g), int(step), C.byref(oob), _dp(terms))
        if want_terms:
            return c, g, int(oob.value), terms
        return c, g, int(oob.value)

    def eval_batch(self, X, step=2, nthreads=0, want_grad=True):
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.n)
        B = X.shape[0]
        cost = np.empty(B)
        grad = np.empty((B, self.n)) if want_grad else None
        oob = C.c_int64(0)
        secs = lib().gto_opt_eval_batch(self.h, _dp(X), B, _dp(cost), _dp(grad) if want_grad else None, int(step),
                                        int(nthreads), C.byref(oob))
        return cost, grad, int(oob.value), secs

    def solve(self, step=2, maxeval=100, maxtime=0.0, xtol_abs=0.0, objective=None, c_objective=None):
        """NLopt LD_MMA solve from the current Dp.  `objective(x, grad) -> float` (optional) replaces
        the oracle objective (numpy views of NLopt's buffers); `c_objective=(fn_ptr, data_ptr)` does the
        same with a native nlopt_func (e.g. the product's gtop_objective and its context)."""
        path = nlopt_path()
        if path is None:
            raise RuntimeError("vendored NLopt binary not available (oracle/_ref/libnlopt.so.0)")
        cb = None
        if objective is not None:
            n = self.n

            def _tramp(nn, xp, gp, _data):
                x = np.ctypeslib.as_array(xp, shape=(n,))
                g = np.ctypeslib.as_array(gp, shape=(n,)) if gp else None
                return float(objective(x, g))

            cb = _NLOPT_FUNC(_tramp)
        x = np.empty(self.n)
        minf = C.c_double(0)
        nev = C.c_int(0)
        fptr, fdata = (C.cast(cb, C.c_void_p), None) if cb else (None, None)
        if c_objective is not None:
            fptr, fdata = c_objective
        r = lib().gto_opt_solve(self.h, int(step), int(maxeval), float(maxtime), float(xtol_abs), path.encode(),
                                fptr, fdata, _dp(x), C.byref(minf), C.byref(nev))
        if r <= -1000:
            raise RuntimeError("oracle solve failed: %d %s" % (r, lib().gto_last_error().decode()))
        return x, minf.value, r, nev.value

    def cost_curve(self):
        n = lib().gto_opt_cost_curve(self.h, None, 0)
        out = np.empty(n)
        lib().gto_opt_cost_curve(self.h, _dp(out), n)
        return out


TRAJ_METRICS = ("time_sum", "length", "jerk", "mean_v", "max_v", "mean_a", "max_a", "acc_cost")


def traj_metrics(coeff, seg_time, want_traj=False):
    """PolynomialTraj evaluation as opti_node runs it (opti_node.cpp:104-144): coeff [m, 18] (getCoefficient), seg_time [m].
    Returns the 8 metrics (TRAJ_METRICS order) and, optionally, the getTraj points [N, 3]."""
    c = np.ascontiguousarray(coeff, dtype=np.float64)
    t = np.ascontiguousarray(seg_time, dtype=np.float64)
    m = t.size
    assert c.shape == (m, 18)
    out = np.empty(8)
    if not want_traj:
        lib().gto_traj_metrics(_dp(c), _dp(t), m, _dp(out), None, 0)
        return out
    cap = int(t.sum() / 0.01) + 8
    traj = np.empty((cap, 3))
    npts = lib().gto_traj_metrics(_dp(c), _dp(t), m, _dp(out), _dp(traj), cap)
    return out, traj[:npts]


def hardware_threads():
    return int(lib().gto_hardware_threads())
