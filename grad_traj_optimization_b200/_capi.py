"""ctypes binding of libgtop_b200.so — the C ABI of include/gtop_c.h.

This is plumbing for tests/ and bench.py; the product is the shared library.  Loading fails
loudly when the library has not been built (there is no Python or CPU fallback).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GTOP_B200_LIB") or os.path.join(_HERE, "libgtop_b200.so")  # env override: kernel A/B builds

# every symbol include/gtop_c.h declares (tests check that the library exports all of them)
SYMBOLS = (
    "gtop_abi_version", "gtop_last_error", "gtop_device_count", "gtop_params_default", "gtop_host_alloc",
    "gtop_host_free", "gtop_sdf_create", "gtop_sdf_destroy", "gtop_sdf_grid", "gtop_sdf_reset",
    "gtop_sdf_reset_range", "gtop_sdf_set_update_range", "gtop_sdf_update_range",
    "gtop_sdf_set_occupancy", "gtop_sdf_set_occupancy_boxes", "gtop_sdf_update_esdf", "gtop_sdf_last_edt_ms", "gtop_sdf_last_update_ms", "gtop_sdf_gather_layout",
    "gtop_sdf_download_distance", "gtop_sdf_upload_distance", "gtop_sdf_download_occupancy",
    "gtop_sdf_upload_occupancy", "gtop_sdf_distance_device_ptr", "gtop_sdf_query", "gtop_sdf_query_dynamic", "gtop_problems_create_dense",
    "gtop_problems_create_from_paths", "gtop_problems_create_from_kino", "gtop_problems_create_from_paths_device",
    "gtop_problems_create_from_kino_device", "gtop_problems_record", "gtop_debug_pack_record", "gtop_traj_metrics_device", "gtop_traj_metrics_host", "gtop_problems_destroy", "gtop_problems_dims", "gtop_problems_set_params",
    "gtop_problems_seg_time", "gtop_problems_x0", "gtop_problems_bounds", "gtop_problems_dense_L",
    "gtop_problems_dense_R", "gtop_problems_Df", "gtop_eval_batch_device", "gtop_eval_batch_host",
    "gtop_problems_take_out_of_map", "gtop_kernel_launch_count", "gtop_coeff_from_x", "gtop_objective_ctx_create",
    "gtop_objective_ctx_set_step", "gtop_objective_ctx_destroy", "gtop_objective_ctx_evals", "gtop_objective",
    "gtop_mma_minimize", "gtop_solve_batch_device", "gtop_solve_batch_host", "gtop_debug_check_sqrt", "gtop_measure_fp64_peak",
)


class GtopParams(C.Structure):
    """Mirror of `struct gtop_params` (include/gtop_c.h)."""
    _fields_ = [(k, C.c_double) for k in (
        "time_limit_1", "time_limit_2", "deltat", "w_smooth", "w_collision", "alpha", "r", "d0",
        "alpha_v", "r_v", "v0", "alpha_a", "r_a", "a0", "bos", "vos", "aos", "mean_v", "mean_a", "init_time")] + [
        (k, C.c_int32) for k in ("algorithm", "samples_per_segment", "vel_acc_penalty", "reserved")]


class GtopError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libgtop_b200.so (once).  Raises if it is missing: build it with __graft_entry__.build()."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GtopError("libgtop_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `make -C grad_traj_optimization_b200/csrc` — there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, dp, i64, i32 = C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.c_int32
    pvp = C.POINTER(C.c_void_p)
    pp = C.POINTER(GtopParams)
    sig = {
        "gtop_abi_version": (C.c_int, []),
        "gtop_last_error": (C.c_char_p, []),
        "gtop_device_count": (C.c_int, []),
        "gtop_params_default": (None, [pp]),
        "gtop_host_alloc": (C.c_int, [pvp, C.c_size_t]),
        "gtop_host_free": (C.c_int, [vp]),
        "gtop_sdf_create": (C.c_int, [dp, C.c_double, dp, C.c_int, pvp]),
        "gtop_sdf_destroy": (C.c_int, [vp]),
        "gtop_sdf_grid": (C.c_int, [vp, C.POINTER(i32)]),
        "gtop_sdf_reset": (C.c_int, [vp]),
        "gtop_sdf_reset_range": (C.c_int, [vp, dp, dp]),
        "gtop_sdf_set_update_range": (C.c_int, [vp, dp, dp]),
        "gtop_sdf_update_range": (C.c_int, [vp, C.POINTER(i32)]),
        "gtop_sdf_set_occupancy": (C.c_int, [vp, dp, i64]),
        "gtop_sdf_set_occupancy_boxes": (C.c_int, [vp, C.POINTER(i32), i64]),
        "gtop_sdf_update_esdf": (C.c_int, [vp]),
        "gtop_sdf_last_edt_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "gtop_sdf_last_update_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "gtop_sdf_gather_layout": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i64)]),
        "gtop_sdf_download_distance": (C.c_int, [vp, dp]),
        "gtop_sdf_upload_distance": (C.c_int, [vp, dp]),
        "gtop_sdf_download_occupancy": (C.c_int, [vp, C.POINTER(i32)]),
        "gtop_sdf_upload_occupancy": (C.c_int, [vp, C.POINTER(i32)]),
        "gtop_sdf_distance_device_ptr": (C.c_int, [vp, pvp]),
        "gtop_sdf_query": (C.c_int, [vp, dp, i64, dp, dp, C.POINTER(i64)]),
        "gtop_sdf_query_dynamic": (C.c_int, [vp, dp, i32, C.c_double, dp, i64, dp, dp, dp]),
        "gtop_problems_create_dense": (C.c_int, [i32, i32, dp, dp, dp, dp, pp, C.c_int, pvp]),
        "gtop_problems_create_from_paths": (C.c_int, [i32, i32, dp, pp, C.c_int, pvp]),
        "gtop_problems_record": (C.c_int, [vp, i32, dp, i64, C.POINTER(i64)]),
        "gtop_debug_pack_record": (C.c_int, [i32, dp, pp, dp, i64, C.POINTER(i32)]),
        "gtop_traj_metrics_device": (C.c_int, [vp, vp, i64, i64, i64, vp, vp]),
        "gtop_traj_metrics_host": (C.c_int, [vp, dp, i64, i64, i64, dp]),
        "gtop_problems_create_from_kino": (C.c_int, [i32, i32, dp, dp, dp, dp, pp, C.c_int, pvp]),
        "gtop_problems_create_from_paths_device": (C.c_int, [i32, i32, dp, pp, C.c_int, pvp]),
        "gtop_problems_create_from_kino_device": (C.c_int, [i32, i32, dp, dp, dp, dp, pp, C.c_int, pvp]),
        "gtop_problems_destroy": (C.c_int, [vp]),
        "gtop_problems_dims": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "gtop_problems_set_params": (C.c_int, [vp, pp]),
        "gtop_problems_seg_time": (C.c_int, [vp, dp]),
        "gtop_problems_x0": (C.c_int, [vp, dp]),
        "gtop_problems_bounds": (C.c_int, [vp, dp, dp]),
        "gtop_problems_dense_L": (C.c_int, [vp, i32, dp]),
        "gtop_problems_dense_R": (C.c_int, [vp, i32, dp]),
        "gtop_problems_Df": (C.c_int, [vp, i32, dp]),
        "gtop_eval_batch_device": (C.c_int, [vp, vp, vp, i64, i64, i64, C.c_int, vp, vp, vp]),
        "gtop_eval_batch_host": (C.c_int, [vp, vp, vp, i64, i64, i64, C.c_int, vp, vp, C.POINTER(i64)]),
        "gtop_problems_take_out_of_map": (C.c_int, [vp, C.POINTER(i64)]),
        "gtop_kernel_launch_count": (i64, []),
        "gtop_coeff_from_x": (C.c_int, [vp, i32, dp, dp]),
        "gtop_objective_ctx_create": (C.c_int, [vp, vp, i32, C.c_int, pvp]),
        "gtop_objective_ctx_set_step": (C.c_int, [vp, C.c_int]),
        "gtop_objective_ctx_destroy": (C.c_int, [vp]),
        "gtop_objective_ctx_evals": (i64, [vp]),
        "gtop_objective": (C.c_double, [C.c_uint, dp, dp, vp]),
        "gtop_mma_minimize": (C.c_int, [C.c_uint, vp, vp, dp, dp, dp, dp, C.c_int, C.POINTER(C.c_int)]),
        "gtop_solve_batch_device": (C.c_int, [vp, vp, vp, i64, i64, i64, C.c_int, C.c_int, vp, vp]),
        "gtop_solve_batch_host": (C.c_int, [vp, vp, vp, i64, i64, i64, C.c_int, C.c_int, vp]),
        "gtop_debug_check_sqrt": (C.c_int, [i64, C.POINTER(i64)]),
        "gtop_measure_fp64_peak": (C.c_int, [C.c_int, dp, dp]),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError when the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise GtopError("gtop error %d: %s" % (rc, lib().gtop_last_error().decode()))
