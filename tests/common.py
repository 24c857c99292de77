"""Shared builders for the tests: fixtures -> oracle objects / GPU handles on identical inputs."""
import numpy as np

from grad_traj_optimization_b200 import fixtures as fx
from oracle import OracleMap, OracleOptimizer

_ALIASES = {"ws": "w_smooth", "wc": "w_collision", "alg": "algorithm", "dt": "deltat"}


def oracle_params(d, **kw):
    out = {}
    for k, v in {**d, **kw}.items():
        k = _ALIASES.get(k, k)
        out[k] = int(v) if k in ("algorithm", "samples_per_segment", "vel_acc_penalty", "exact_gradient") else float(v)
    return out


def oracle_map_from_cfg(cfg):
    om = OracleMap(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
    om.reset()
    om.set_occupancy(cfg["obstacles"])
    om.update_esdf()
    return om


def oracle_problem(cfg_or_params, om, waypoints, **kw):
    params = cfg_or_params["params"] if "params" in cfg_or_params else cfg_or_params
    opt = OracleOptimizer(oracle_params(params, **kw), om)
    opt.set_path(waypoints)
    return opt


def rel_cost_err(c, c_ref):
    return np.abs(c - c_ref) / np.abs(c_ref)


def rel_grad_err(g, g_ref):
    """||g - g_ref||_inf / ||g_ref||_inf per trajectory (SURVEY §8d parity definition)."""
    g, g_ref = np.atleast_2d(g), np.atleast_2d(g_ref)
    return np.abs(g - g_ref).max(axis=1) / np.abs(g_ref).max(axis=1)


_map_a_cache = {}


def map_a_oracle():
    """Map A (20x20x5 m @0.1 m, 120 random pillars) built by the oracle; cached per process."""
    if "om" not in _map_a_cache:
        om = OracleMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
        om.reset()
        om.set_occupancy(fx.map_a_obstacles())
        om.update_esdf()
        _map_a_cache["om"] = om
        _map_a_cache["dist"] = om.distance()
    return _map_a_cache["om"], _map_a_cache["dist"]


def map_a_batch(m, n_problems, n_starts, first_problem=0):
    """waypoint sets + multi-start X for Map A: returns (waypoints[np,m+1,3], X[np*n_starts, n])."""
    _, dist = map_a_oracle()
    wps, Xs = [], []
    for p in range(n_problems):
        w = fx.random_waypoints(first_problem + p, m, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"])
        x0, lb, ub = fx.straight_line_x0(w, fx.PARAMS_OPTI["bos"], fx.PARAMS_OPTI["vos"], fx.PARAMS_OPTI["aos"])
        wps.append(w)
        Xs.append(fx.multi_starts(x0, lb, ub, n_starts, seed=7 + first_problem + p))
    return np.asarray(wps), np.concatenate(Xs, axis=0)
