"""The oracle against the REFERENCE ITSELF (no GPU).

oracle/_ref/libgtop_reference.so is the reference's src/sdf_map.cpp, src/qp_generator.cpp and
src/grad_traj_optimizer.cpp compiled unmodified against the stand-in headers of oracle/shim/ (`make -C oracle
reference`, only where /root/reference exists; the built library travels with the repo snapshot).  Two layers:

  * direct: the oracle restatement and the reference build on the same inputs — distance fields and trilinear
    queries bit for bit (everything in sdf_map.cpp is element-wise arithmetic), set-up matrices / objective /
    MMA iterates to rounding level (1e-12; in practice they come out identical);
  * golden: tests/golden/reference_run.npz holds outputs of the reference build (make_reference_golden.py), so
    the same pins hold where the library cannot be rebuilt; the CUDA path is checked against them in
    tests/test_gpu_reference_golden.py.
"""
import hashlib
import os

import numpy as np
import pytest

from common import map_a_batch, map_a_oracle, oracle_map_from_cfg, oracle_problem, rel_cost_err, rel_grad_err
from grad_traj_optimization_b200 import fixtures as fx
from oracle import OracleMap, nlopt_path
from oracle import ref_binding as rb

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_run.npz")
needs_ref = pytest.mark.skipif(not rb.available(), reason="oracle/_ref/libgtop_reference.so not built (needs /root/reference)")


def ref_map_from_cfg(cfg):
    rm = rb.ReferenceMap(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
    rm.reset(), rm.set_occupancy(cfg["obstacles"]), rm.update_esdf()
    return rm


def ref_problem(cfg, waypoints, dist=None, params=None):
    ro = rb.ReferenceOptimizer(params or cfg["params"])
    ro.init_map(cfg["map_size"], cfg["map_origin"], cfg["resolution"])
    if dist is None:
        ro.update_map(cfg["obstacles"])
    else:
        ro.set_distance(dist)
    ro.set_path(waypoints)
    return ro


@pytest.fixture(scope="module", params=["text_input", "opti_node"])
def case(request):
    cfg = fx.load_cfg(request.param)
    om = oracle_map_from_cfg(cfg)
    return request.param, cfg, om, oracle_problem(cfg, om, cfg["waypoints"])


# ---- direct comparisons ------------------------------------------------------------------------------------
@needs_ref
def test_edt_bit_equal_to_reference_sdf_map(case):
    """updateESDF3d / fillESDF / setOccupancy / resetBuffer (sdf_map.cpp:26-99, :266-368)."""
    name, cfg, om, _ = case
    rm = ref_map_from_cfg(cfg)
    assert rm.grid == om.grid
    assert np.array_equal(rm.occupancy(), om.occupancy())
    assert np.array_equal(rm.distance(), om.distance())


@needs_ref
def test_edt_bit_equal_on_map_a_and_small_grids():
    om, dist = map_a_oracle()
    rm = rb.ReferenceMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
    rm.reset(), rm.set_occupancy(fx.map_a_obstacles()), rm.update_esdf()
    assert np.array_equal(rm.distance(), dist)
    rng = np.random.default_rng(3)
    for size, res in (((1.0, 0.7, 0.5), 0.1), ((0.9, 0.3, 1.7), 0.05), ((3.0, 0.2, 0.2), 0.1), ((0.4, 0.4, 0.4), 0.2)):
        for fill in (0.0, 0.02, 0.5):
            a = OracleMap((-0.3, 0.1, 0.0), res, size)
            b = rb.ReferenceMap((-0.3, 0.1, 0.0), res, size)
            pts = np.array([-0.3, 0.1, 0.0]) + np.array(size) * rng.random((int(fill * a.nvox) + (fill > 0), 3))
            for mp in (a, b):
                mp.reset(), mp.set_occupancy(pts), mp.update_esdf()
            assert np.array_equal(a.occupancy(), b.occupancy())
            assert np.array_equal(a.distance(), b.distance()), (size, res, fill)


@needs_ref
def test_windowed_updates_bit_equal(case):
    """setUpdateRange / resetBuffer(min, max) (sdf_map.cpp:28-53, :244-264, :366-367) through a sequence of updates."""
    name, cfg, _, _ = case
    wp = np.asarray(cfg["waypoints"], dtype=np.float64)
    lo, hi = wp.min(axis=0) - 0.7, wp.max(axis=0) + 0.4
    extra = wp[1:-1] + np.array([0.35, -0.2, 0.1])
    maps = [oracle_map_from_cfg(cfg), ref_map_from_cfg(cfg)]
    for mp in maps:
        mp.reset_range(lo, hi), mp.set_occupancy(extra), mp.set_update_range(lo, hi)
    assert all(np.array_equal(a, b) for a, b in zip(maps[0].update_range(), maps[1].update_range()))
    for mp in maps:
        mp.update_esdf()
    assert np.array_equal(maps[0].distance(), maps[1].distance())
    assert all(np.array_equal(a, b) for a, b in zip(maps[0].update_range(), maps[1].update_range()))  # window reset
    for mp in maps:  # a second, partly out-of-map window, then a full rebuild
        mp.set_update_range(lo - 30.0, 0.5 * (lo + hi)), mp.set_occupancy(extra + 0.3), mp.update_esdf()
    assert np.array_equal(maps[0].distance(), maps[1].distance())
    for mp in maps:
        mp.update_esdf()
    assert np.array_equal(maps[0].distance(), maps[1].distance())


@needs_ref
def test_trilinear_query_bit_equal(case):
    """getDistWithGradTrilinear (sdf_map.cpp:185-242) incl. the 1e-4 in-map margin, faces and outside points."""
    from golden.make_reference_golden import query_points
    name, cfg, om, _ = case
    rm = ref_map_from_cfg(cfg)
    q = query_points(cfg, n=4096, seed=23)
    d0, g0, oob = om.query(q)
    d1, g1 = rm.query(q)
    assert oob == int((d1 == -1.0).sum()) and oob > 100
    assert np.array_equal(d0, d1) and np.array_equal(g0, g1)


@needs_ref
def test_setup_matrices_and_objective_match_reference(case):
    """setPath (grad_traj_optimizer.cpp:67-110) through qp_generator.cpp, then costFunc (:554-562) in both steps."""
    name, cfg, om, opt = case
    ro = ref_problem(cfg, cfg["waypoints"])
    assert (ro.m, ro.n) == (opt.m, opt.n)
    assert np.array_equal(ro.seg_time(), opt.seg_time())
    for a, b in ((ro.L(), opt.L()), (ro.R(), opt.R()), (ro.Df(), opt.Df()), (ro.x0(), opt.x0())):
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())
    lb, ub = opt.bounds()
    X = fx.multi_starts(opt.x0(), lb, ub, 24, seed=17)
    for step in (1, 2):
        c, g = ro.eval_batch(X, step)
        cr, gr, _, _ = opt.eval_batch(X, step)
        assert rel_cost_err(cr, c).max() <= 1e-12 and rel_grad_err(gr, g).max() <= 1e-12
    assert np.abs(ro.coeff_from_x(X[3]) - opt.coeff_from_x(X[3])).max() <= 1e-12 * np.abs(opt.coeff_from_x(X[3])).max()


@needs_ref
@pytest.mark.parametrize("m", [6, 10])
def test_objective_matches_reference_on_map_a(m):
    """random.launch-style map, random waypoint sets, multi-starts: BASELINE configs 2 / 3 shapes."""
    om, dist = map_a_oracle()
    wps, X = map_a_batch(m, 3, 8)
    cfg = dict(map_size=fx.MAP_A["map_size"], map_origin=fx.MAP_A["map_origin"], resolution=fx.MAP_A["resolution"])
    for p in range(3):
        ro = ref_problem(cfg, wps[p], dist=dist, params=fx.PARAMS_OPTI)
        opt = oracle_problem(fx.PARAMS_OPTI, om, wps[p])
        c, g = ro.eval_batch(X[8 * p:8 * p + 8], 2)
        cr, gr, _, _ = opt.eval_batch(X[8 * p:8 * p + 8], 2)
        assert rel_cost_err(cr, c).max() <= 1e-12 and rel_grad_err(gr, g).max() <= 1e-12


@needs_ref
def test_out_of_map_samples_follow_the_ruling():
    """the reference leaves `grad` unwritten for out-of-map samples; the shim's vectors start at zero, the
    oracle sets zero.  A start pushed above the map must agree."""
    cfg = fx.load_cfg("text_input")
    om = oracle_map_from_cfg(cfg)
    opt = oracle_problem(cfg, om, cfg["waypoints"])
    ro = ref_problem(cfg, cfg["waypoints"])
    x = opt.x0().copy()
    P = opt.num_dp
    x[2 * P:2 * P + 9:3] += 4.0  # z of the first three interior waypoints: above the 5 m ceiling
    c, g = ro.eval_batch(x[None], 2)
    cr, gr, oob = opt.eval(x, 2)
    assert oob > 0
    assert rel_cost_err(np.array([cr]), c).max() <= 1e-12 and rel_grad_err(gr, g).max() <= 1e-12


@needs_ref
@pytest.mark.skipif(nlopt_path() is None, reason="vendored NLopt binary not available")
def test_mma_solve_matches_reference(case):
    """optimizeTrajectory(OPT_SECOND_STEP) (:128-243) stopped after 40 evaluations, through the same NLopt 2.5.0 binary."""
    name, cfg, om, opt = case
    ro = ref_problem(cfg, cfg["waypoints"])
    x_start = ro.x0()
    x_ref, ev = ro.optimize(2, 40)
    x, minf, rc, nev = oracle_problem(cfg, om, cfg["waypoints"]).solve(step=2, maxeval=40)
    assert ev == nev == 40
    assert np.abs(x - x_ref).max() <= 1e-9 * max(1.0, np.abs(x_ref).max())
    assert np.abs(x_ref - x_start).max() > 1e-3  # it did move


@needs_ref
def test_kino_path_setup_matches_reference():
    """setKinoPath (:35-65) / PolyKinoGeneration (qp_generator.cpp:23-154)."""
    cfg = fx.load_cfg("opti_node")
    om = oracle_map_from_cfg(cfg)
    wp = np.asarray(cfg["waypoints"], dtype=np.float64)
    rng = np.random.default_rng(9)
    vel, acc = rng.normal(0, 0.8, wp.shape), rng.normal(0, 0.5, wp.shape)
    T = 0.6 + rng.random(len(wp) - 1)
    ro = rb.ReferenceOptimizer(cfg["params"])
    ro.init_map(cfg["map_size"], cfg["map_origin"], cfg["resolution"]), ro.update_map(cfg["obstacles"])
    ro.set_kino_path(wp, vel, acc, T)
    from common import oracle_params
    from oracle import OracleOptimizer
    opt = OracleOptimizer(oracle_params(cfg["params"]), om)
    opt.set_kino_path(wp, vel, acc, T)
    for a, b in ((ro.L(), opt.L()), (ro.R(), opt.R()), (ro.Df(), opt.Df()), (ro.x0(), opt.x0()), (ro.coeff(), opt.coeff())):
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())
    c, g = ro.eval_batch(opt.x0()[None], 2)
    cr, gr, _ = opt.eval(opt.x0(), 2)
    assert rel_cost_err(np.array([cr]), c).max() <= 1e-12 and rel_grad_err(gr, g).max() <= 1e-12


# ---- golden vectors of the reference build (these run wherever the repo is) -----------------------------------
def test_oracle_reproduces_reference_golden(case):
    name, cfg, om, opt = case
    gold = np.load(GOLD)
    d = om.distance()
    assert tuple(gold[name + "_grid"]) == d.shape
    assert hashlib.sha256(d.tobytes()).digest() == gold[name + "_dist_sha256"].tobytes()
    assert np.array_equal(d[::7, ::5, ::3], gold[name + "_dist_sub"]) and d.max() == gold[name + "_dist_max"]
    qd, qg, _ = om.query(gold[name + "_query_pos"])
    assert np.array_equal(qd, gold[name + "_query_dist"]) and np.array_equal(qg, gold[name + "_query_grad"])
    assert np.array_equal(opt.seg_time(), gold[name + "_seg_time"])
    for key, val in (("_L", opt.L()), ("_R", opt.R()), ("_Df", opt.Df()), ("_x0", opt.x0())):
        assert np.abs(val - gold[name + key]).max() <= 1e-12 * max(1.0, np.abs(val).max())
    for step in (1, 2):
        c, g, oob, _ = opt.eval_batch(gold[name + "_X"], step)
        assert rel_cost_err(c, gold["%s_cost%d" % (name, step)]).max() <= 1e-12
        assert rel_grad_err(g, gold["%s_grad%d" % (name, step)]).max() <= 1e-12
    if nlopt_path() is not None:
        x, _, _, nev = oracle_problem(cfg, om, cfg["waypoints"]).solve(step=2, maxeval=40)
        assert nev == int(gold[name + "_solve40_evals"]) == 40
        assert np.abs(x - gold[name + "_solve40_x"]).max() <= 1e-9 * max(1.0, np.abs(x).max())
