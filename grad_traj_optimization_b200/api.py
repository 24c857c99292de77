"""Thin Python handles over the C ABI (include/gtop_c.h) — used by tests/ and bench.py.

Names follow the reference's classes: `SdfMap` ~ SDFMap (sdf_map.h), `ProblemSet` ~ the matrices a
GradTrajOptimizer holds after setPath (grad_traj_optimizer.h:44-57), `Objective` ~ the NLopt
callback GradTrajOptimizer::costFunc (grad_traj_optimizer.cpp:554-562).  Host data is numpy;
device-resident batches are torch CUDA tensors whose data_ptr() is handed to the library.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import GtopParams, check, lib

OPT_INITIAL_TRY, OPT_FIRST_STEP, OPT_SECOND_STEP = 0, 1, 2

# launch-file spellings (grad_traj_optimizer.cpp:5-32) -> struct fields
_PARAM_ALIASES = {"ws": "w_smooth", "wc": "w_collision", "alg": "algorithm", "dt": "deltat"}


def make_params(d=None, **kw):
    """gtop_params from a dict of launch-file style names (ws, wc, alpha, ...) or struct field names."""
    p = GtopParams()
    lib().gtop_params_default(C.byref(p))
    merged = dict(d or {})
    merged.update(kw)
    names = {f[0] for f in GtopParams._fields_}
    for k, v in merged.items():
        k = _PARAM_ALIASES.get(k, k)
        if k in names:
            setattr(p, k, int(v) if k in ("algorithm", "samples_per_segment", "vel_acc_penalty") else float(v))
    return p


def params_to_dict(p):
    return {f[0]: getattr(p, f[0]) for f in GtopParams._fields_ if f[0] != "reserved"}


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


class SdfMap:
    """Voxel map + ESDF resident on one GPU (SDFMap, sdf_map.cpp)."""

    def __init__(self, origin, resolution, size, device=0):
        self.h = C.c_void_p()
        o, s = _f64(origin), _f64(size)
        check(lib().gtop_sdf_create(_dp(o), float(resolution), _dp(s), int(device), C.byref(self.h)))
        g = (C.c_int32 * 3)()
        check(lib().gtop_sdf_grid(self.h, g))
        self.grid = tuple(int(v) for v in g)
        self.origin, self.resolution, self.size, self.device = o.copy(), float(resolution), s.copy(), int(device)

    def close(self):
        if getattr(self, "h", None) and lib is not None:  # (module globals are gone at interpreter shutdown)
            lib().gtop_sdf_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def nvox(self):
        return self.grid[0] * self.grid[1] * self.grid[2]

    def reset(self):
        check(lib().gtop_sdf_reset(self.h))

    def reset_range(self, min_pos, max_pos):
        """SDFMap::resetBuffer(min_pos, max_pos) (sdf_map.cpp:28-53)."""
        a, b = _f64(min_pos, (3,)), _f64(max_pos, (3,))
        check(lib().gtop_sdf_reset_range(self.h, _dp(a), _dp(b)))

    def set_update_range(self, min_pos, max_pos):
        """SDFMap::setUpdateRange (sdf_map.cpp:244-264): window of the next update_esdf only."""
        a, b = _f64(min_pos, (3,)), _f64(max_pos, (3,))
        check(lib().gtop_sdf_set_update_range(self.h, _dp(a), _dp(b)))

    def update_range(self):
        out = np.zeros(6, dtype=np.int32)
        check(lib().gtop_sdf_update_range(self.h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out[:3].copy(), out[3:].copy()

    def set_occupancy(self, pts):
        p = _f64(pts, (-1, 3))
        check(lib().gtop_sdf_set_occupancy(self.h, _dp(p), p.shape[0]))

    def set_occupancy_boxes(self, lo_hi):
        b = np.ascontiguousarray(lo_hi, dtype=np.int32).reshape(-1, 6)
        check(lib().gtop_sdf_set_occupancy_boxes(self.h, b.ctypes.data_as(C.POINTER(C.c_int32)), b.shape[0]))

    def update_esdf(self):
        check(lib().gtop_sdf_update_esdf(self.h))

    def update(self, obstacle_points):
        """GradTrajOptimizer::updateSDFMap (grad_traj_optimizer.cpp:117-126)."""
        self.reset()
        self.set_occupancy(obstacle_points)
        self.update_esdf()

    def last_edt_ms(self):
        ms = (C.c_float * 4)()
        check(lib().gtop_sdf_last_edt_ms(self.h, ms))
        return [float(v) for v in ms]

    def last_update_ms(self):
        """z, y, x passes, their sum, rebuild of the gather copy, update start -> ready for the evaluation kernels."""
        ms = (C.c_float * 6)()
        check(lib().gtop_sdf_last_update_ms(self.h, ms))
        return [float(v) for v in ms]

    def gather_layout(self):
        """(has_quad_copy, resident device bytes): see gtop_sdf_gather_layout."""
        q, b = C.c_int32(0), C.c_int64(0)
        check(lib().gtop_sdf_gather_layout(self.h, C.byref(q), C.byref(b)))
        return bool(q.value), int(b.value)

    def distance(self):
        out = np.empty(self.nvox, dtype=np.float64)
        check(lib().gtop_sdf_download_distance(self.h, _dp(out)))
        return out.reshape(self.grid)

    def upload_distance(self, d):
        d = _f64(d, (-1,))
        assert d.size == self.nvox
        check(lib().gtop_sdf_upload_distance(self.h, _dp(d)))

    def occupancy(self):
        out = np.empty(self.nvox, dtype=np.int32)
        check(lib().gtop_sdf_download_occupancy(self.h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out.reshape(self.grid)

    def upload_occupancy(self, occ):
        o = np.ascontiguousarray(occ, dtype=np.int32).reshape(-1)
        assert o.size == self.nvox
        check(lib().gtop_sdf_upload_occupancy(self.h, o.ctypes.data_as(C.POINTER(C.c_int32))))

    def query(self, pos):
        """Batched getDistWithGradTrilinear: (dist[n], grad[n,3], n_out_of_map)."""
        p = _f64(pos, (-1, 3))
        dist = np.empty(p.shape[0])
        grad = np.empty((p.shape[0], 3))
        oob = C.c_int64(0)
        check(lib().gtop_sdf_query(self.h, _dp(p), p.shape[0], _dp(dist), _dp(grad), C.byref(oob)))
        return dist, grad, int(oob.value)


    def query_dynamic(self, pos, boxes, time):
        """Batched EDTEnvironment::evaluateEDTWithGrad / evaluateCoarseEDT (edt_environment.cpp:75-137): static field
        min'ed with constant-velocity boxes [nbox, 9] = (centre at t=0, velocity, edge lengths) at `time`.
        Returns (dist[n], grad[n,3], coarse[n])."""
        p = _f64(pos, (-1, 3))
        b = _f64(boxes, (-1, 9))
        dist, grad, coarse = np.empty(p.shape[0]), np.empty((p.shape[0], 3)), np.empty(p.shape[0])
        check(lib().gtop_sdf_query_dynamic(self.h, _dp(b), b.shape[0], float(time), _dp(p), p.shape[0], _dp(dist),
                                           _dp(grad), _dp(coarse)))
        return dist, grad, coarse


class ProblemSet:
    """A set of trajectory problems with the same segment count m, resident on one GPU."""

    def __init__(self, handle, params, device):
        self.h, self.params, self.device = handle, params, device
        np_, m, n = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib().gtop_problems_dims(self.h, C.byref(np_), C.byref(m), C.byref(n)))
        self.num_problems, self.m, self.n = np_.value, m.value, n.value
        self.num_dp = self.n // 3

    @classmethod
    def from_dense(cls, seg_time, L, R, Df, params, device=0):
        """Drop-in path: L, R (column-major values as 2-D/3-D numpy arrays indexed [.., row, col]),
        Df [.., 3, 6], seg_time [.., m] as left by the reference's setPath."""
        seg_time = _f64(seg_time)
        seg_time = seg_time.reshape(-1, seg_time.shape[-1])
        np_, m = seg_time.shape
        Lc = np.ascontiguousarray(np.swapaxes(_f64(L).reshape(np_, 6 * m, 3 * m + 3), 1, 2))  # -> column-major
        Rc = np.ascontiguousarray(np.swapaxes(_f64(R).reshape(np_, 3 * m + 3, 3 * m + 3), 1, 2))
        Dc = np.ascontiguousarray(np.swapaxes(_f64(Df).reshape(np_, 3, 6), 1, 2))
        p = params if isinstance(params, GtopParams) else make_params(params)
        h = C.c_void_p()
        check(lib().gtop_problems_create_dense(np_, m, _dp(seg_time), _dp(Lc), _dp(Rc), _dp(Dc), C.byref(p), int(device),
                                               C.byref(h)))
        return cls(h, p, int(device))

    @classmethod
    def from_paths(cls, waypoints, params, device=0, setup="host"):
        """Native path: waypoints [num_problems, m+1, 3] (setPath, grad_traj_optimizer.cpp:67-110).
        setup="device" builds the per-problem records with the set-up kernel instead of the host threads."""
        w = _f64(waypoints)
        if w.ndim == 2:
            w = w[None]
        np_, m = w.shape[0], w.shape[1] - 1
        p = params if isinstance(params, GtopParams) else make_params(params)
        h = C.c_void_p()
        create = {"host": lib().gtop_problems_create_from_paths, "device": lib().gtop_problems_create_from_paths_device}[setup]
        check(create(np_, m, _dp(w), C.byref(p), int(device), C.byref(h)))
        return cls(h, p, int(device))

    @classmethod
    def from_kino(cls, pos, vel, acc, time, params, device=0, setup="host"):
        """Front-end hand-off (setKinoPath, grad_traj_optimizer.cpp:35-65): pos/vel/acc [np, m+1, 3], time [np, m]."""
        t = _f64(time)
        t = t.reshape(-1, t.shape[-1])
        np_, m = t.shape
        P, V, A = (_f64(a, (np_, m + 1, 3)) for a in (pos, vel, acc))
        p = params if isinstance(params, GtopParams) else make_params(params)
        h = C.c_void_p()
        create = {"host": lib().gtop_problems_create_from_kino, "device": lib().gtop_problems_create_from_kino_device}[setup]
        check(create(np_, m, _dp(P), _dp(V), _dp(A), _dp(t), C.byref(p), int(device), C.byref(h)))
        return cls(h, p, int(device))

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib().gtop_problems_destroy(self.h)
            self.h = None

    __del__ = close

    METRICS = ("time_sum", "length", "jerk", "mean_v", "max_v", "mean_a", "max_a", "acc_cost")

    def traj_metrics_host(self, X, traj_offset=0, starts_per_problem=1):
        """PolynomialTraj evaluation of a batch of variable vectors (opti_node.cpp:104-144): [B, 8] in METRICS order."""
        X = _f64(X, (-1, self.n))
        out = np.empty((X.shape[0], 8))
        check(lib().gtop_traj_metrics_host(self.h, _dp(X), X.shape[0], int(traj_offset), int(starts_per_problem), _dp(out)))
        return out

    def traj_metrics_device(self, X, out, traj_offset=0, starts_per_problem=1, stream=None):
        """Same with device-resident torch tensors X [B, n] and out [B, 8] (float64, contiguous)."""
        check(lib().gtop_traj_metrics_device(self.h, X.data_ptr(), X.shape[0], int(traj_offset), int(starts_per_problem),
                                             out.data_ptr(), stream))
        return out

    def record(self, problem=0):
        """The packed device record of one problem (inspection / tests)."""
        n = C.c_int64(0)
        check(lib().gtop_problems_record(self.h, int(problem), None, 0, C.byref(n)))
        out = np.empty(n.value)
        check(lib().gtop_problems_record(self.h, int(problem), _dp(out), n.value, None))
        return out

    def layout_offsets(self):
        """Offsets (in doubles) of the first sections of a record: dt, ns, Linv, tt, Rc (csrc/gtop_types.cuh)."""
        m, ns = self.m, int(self.params.samples_per_segment)
        o = {"dt": 0, "ns": m, "Linv": 2 * m, "tt": 2 * m + 36 * m}
        o["Rc"] = o["tt"] + 5 * m * ns
        return o

    def set_params(self, params):
        p = params if isinstance(params, GtopParams) else make_params(params)
        check(lib().gtop_problems_set_params(self.h, C.byref(p)))
        self.params = p

    def seg_time(self):
        out = np.empty((self.num_problems, self.m))
        check(lib().gtop_problems_seg_time(self.h, _dp(out)))
        return out

    def x0(self):
        out = np.empty((self.num_problems, self.n))
        check(lib().gtop_problems_x0(self.h, _dp(out)))
        return out

    def bounds(self):
        lb, ub = np.empty((self.num_problems, self.n)), np.empty((self.num_problems, self.n))
        check(lib().gtop_problems_bounds(self.h, _dp(lb), _dp(ub)))
        return lb, ub

    def dense_L(self, i=0):
        out = np.empty((3 * self.m + 3, 6 * self.m))  # column-major storage
        check(lib().gtop_problems_dense_L(self.h, int(i), _dp(out)))
        return out.T.copy()

    def dense_R(self, i=0):
        out = np.empty((3 * self.m + 3, 3 * self.m + 3))
        check(lib().gtop_problems_dense_R(self.h, int(i), _dp(out)))
        return out.T.copy()

    def Df(self, i=0):
        out = np.empty((6, 3))
        check(lib().gtop_problems_Df(self.h, int(i), _dp(out)))
        return out.T.copy()

    def coeff_from_x(self, x, i=0):
        x = _f64(x, (-1,))
        out = np.empty((self.m, 18))
        check(lib().gtop_coeff_from_x(self.h, int(i), _dp(x), _dp(out)))
        return out

    def take_out_of_map(self):
        n = C.c_int64(0)
        check(lib().gtop_problems_take_out_of_map(self.h, C.byref(n)))
        return int(n.value)

    # ---- evaluation ---------------------------------------------------------------------------
    def eval_host(self, sdf, X, step=OPT_SECOND_STEP, traj_offset=0, starts_per_problem=1, cost=None, grad=None):
        """getCostAndGradient for a batch held in host memory (numpy, or pinned torch CPU tensors):
        copies X in, evaluates on the GPU, copies cost/grad out.  Returns (cost[B], grad[B,n], oob)."""
        if hasattr(X, "data_ptr"):  # torch CPU tensor
            B = X.shape[0]
            xp, cp, gp = X.data_ptr(), cost.data_ptr(), grad.data_ptr()
        else:
            X = _f64(X, (-1, self.n))
            B = X.shape[0]
            cost = np.empty(B) if cost is None else cost
            grad = np.empty((B, self.n)) if grad is None else grad
            xp, cp, gp = X.ctypes.data, cost.ctypes.data, grad.ctypes.data
        oob = C.c_int64(0)
        check(lib().gtop_eval_batch_host(sdf.h, self.h, xp, B, int(traj_offset), int(starts_per_problem), int(step), cp, gp,
                                         C.byref(oob)))
        return cost, grad, int(oob.value)

    def eval_device(self, sdf, X, cost, grad, step=OPT_SECOND_STEP, traj_offset=0, starts_per_problem=1, stream=None):
        """Enqueue one evaluation of a device-resident batch (torch CUDA float64 tensors)."""
        B = X.shape[0]
        check(lib().gtop_eval_batch_device(sdf.h, self.h, X.data_ptr(), B, int(traj_offset), int(starts_per_problem),
                                           int(step), cost.data_ptr(), grad.data_ptr(), stream))


    def solve_host(self, sdf, X, step=OPT_SECOND_STEP, maxeval=100, traj_offset=0, starts_per_problem=1, minf=None):
        """Batched lock-step MMA (the optimiser loop of optimizeTrajectory) on the GPU: returns (best X, min f).
        A numpy X is copied; a (pinned) torch CPU tensor is solved IN PLACE and `minf` must be a torch CPU tensor too."""
        if hasattr(X, "data_ptr"):
            check(lib().gtop_solve_batch_host(sdf.h, self.h, X.data_ptr(), X.shape[0], int(traj_offset), int(starts_per_problem),
                                              int(step), int(maxeval), minf.data_ptr()))
            return X, minf
        X = np.array(X, dtype=np.float64, order="C").reshape(-1, self.n)
        minf = np.empty(X.shape[0])
        check(lib().gtop_solve_batch_host(sdf.h, self.h, _dp(X), X.shape[0], int(traj_offset), int(starts_per_problem),
                                          int(step), int(maxeval), _dp(minf)))
        return X, minf

    def solve_device(self, sdf, X, minf, step=OPT_SECOND_STEP, maxeval=100, traj_offset=0, starts_per_problem=1,
                     stream=None):
        """Enqueue the batched solve on device-resident starts (torch CUDA float64 tensors, X updated in place)."""
        check(lib().gtop_solve_batch_device(sdf.h, self.h, X.data_ptr(), X.shape[0], int(traj_offset),
                                            int(starts_per_problem), int(step), int(maxeval),
                                            minf.data_ptr() if minf is not None else None, stream))


def mma_minimize(fn_ptr, fn_data, lb, ub, x0, maxeval):
    """gtop_mma_minimize: host MMA loop around a native nlopt_func (C function pointer + data pointer)."""
    lb, ub = np.ascontiguousarray(lb, dtype=np.float64), np.ascontiguousarray(ub, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, order="C")
    minf, nev = C.c_double(0), C.c_int(0)
    check(lib().gtop_mma_minimize(x.size, fn_ptr, fn_data, _dp(lb), _dp(ub), _dp(x),
                                  C.cast(C.byref(minf), C.POINTER(C.c_double)), int(maxeval), C.byref(nev)))
    return x, minf.value, nev.value


class Objective:
    """`this` + costFunc: an NLopt-compatible callback bound to one problem of a ProblemSet."""

    def __init__(self, sdf, problems, problem=0, step=OPT_SECOND_STEP):
        self.sdf, self.problems = sdf, problems  # keep alive
        self.h = C.c_void_p()
        check(lib().gtop_objective_ctx_create(sdf.h, problems.h, int(problem), int(step), C.byref(self.h)))
        self.n = problems.n

    def close(self):
        if getattr(self, "h", None):
            lib().gtop_objective_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    def set_step(self, step):
        check(lib().gtop_objective_ctx_set_step(self.h, int(step)))

    @property
    def evals(self):
        return int(lib().gtop_objective_ctx_evals(self.h))

    @property
    def c_callback(self):
        """(function pointer, data pointer) to hand to nlopt_set_min_objective."""
        return C.cast(lib().gtop_objective, C.c_void_p), self.h

    def __call__(self, x, grad=None):
        x = _f64(x, (-1,))
        g = np.empty(self.n) if grad is None else grad
        c = lib().gtop_objective(self.n, _dp(x), _dp(g), self.h)
        if not np.isfinite(c):
            raise _capi.GtopError(lib().gtop_last_error().decode())
        return (c, g) if grad is None else c
