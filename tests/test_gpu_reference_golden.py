"""The CUDA path against outputs of the REFERENCE ITSELF (GPU).

tests/golden/reference_run.npz holds what the reference's own src/sdf_map.cpp, src/qp_generator.cpp and
src/grad_traj_optimizer.cpp — compiled unmodified against the stand-in headers of oracle/shim/ — compute on its
two fixtures (tests/golden/make_reference_golden.py).  Here the product runs the same inputs through the C ABI:
distance field and trilinear query bit for bit, cost / gradient within the 1e-9 gate, the on-device MMA loop within
1e-4 m of the reference's optimizeTrajectory."""
import hashlib
import os

import numpy as np
import pytest

from common import rel_cost_err, rel_grad_err
from grad_traj_optimization_b200 import fixtures as fx

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_run.npz")
TOL = 1e-9


@pytest.fixture(scope="module", params=["text_input", "opti_node"])
def case(request, gtop_lib):
    from grad_traj_optimization_b200 import api
    cfg = fx.load_cfg(request.param)
    gm = api.SdfMap(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
    gm.update(cfg["obstacles"])  # resetBuffer + setOccupancy + updateESDF3d on the device
    return request.param, cfg, gm, np.load(GOLD), api


def test_gpu_edt_equals_reference_run(case):
    """updateSDFMap (grad_traj_optimizer.cpp:117-126 -> sdf_map.cpp:26-99, :310-368)."""
    name, cfg, gm, gold, api = case
    d = gm.distance()
    assert d.shape == tuple(gold[name + "_grid"])
    assert hashlib.sha256(d.tobytes()).digest() == gold[name + "_dist_sha256"].tobytes()
    assert np.array_equal(d[::7, ::5, ::3], gold[name + "_dist_sub"])


def test_gpu_trilinear_query_equals_reference_run(case):
    """getDistWithGradTrilinear (sdf_map.cpp:185-242): interior, faces, out of map."""
    name, cfg, gm, gold, api = case
    dist, grad, oob = gm.query(gold[name + "_query_pos"])
    assert np.array_equal(dist, gold[name + "_query_dist"]) and np.array_equal(grad, gold[name + "_query_grad"])
    assert oob == int((gold[name + "_query_dist"] == -1.0).sum())


@pytest.mark.parametrize("setup", ["host", "device"])
@pytest.mark.parametrize("step", [1, 2])
def test_gpu_objective_equals_reference_run(case, step, setup):
    """costFunc (grad_traj_optimizer.cpp:554-562) on 16 multi-starts; set-up by the product (host / device)."""
    name, cfg, gm, gold, api = case
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]), setup=setup)
    assert np.array_equal(ps.seg_time()[0], gold[name + "_seg_time"]) or setup == "device"
    assert np.abs(ps.x0()[0] - gold[name + "_x0"]).max() <= 1e-12 * np.abs(gold[name + "_x0"]).max()
    X = gold[name + "_X"]
    for spp in (X.shape[0], 1):  # multi-start kernel (k_eval_mma) and the general kernel (k_eval_traj)
        Xs = X if spp > 1 else X[:1]
        c, g, oob = ps.eval_host(gm, Xs, step=step, starts_per_problem=spp)
        assert rel_cost_err(c, gold["%s_cost%d" % (name, step)][:len(c)]).max() <= TOL
        assert rel_grad_err(g, gold["%s_grad%d" % (name, step)][:len(c)]).max() <= TOL


def test_gpu_setup_matrices_equal_reference_run(case):
    """setPath -> PolyQPGeneration / StackOptiDep (qp_generator.cpp:156-451): L, R, Df from the product's block-wise set-up."""
    name, cfg, gm, gold, api = case
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
    for key, val in (("_L", ps.dense_L()), ("_R", ps.dense_R()), ("_Df", ps.Df())):
        assert np.abs(val - gold[name + key]).max() <= 1e-12 * np.abs(gold[name + key]).max()


def test_gpu_mma_solve_close_to_reference_run(case):
    """optimizeTrajectory(OPT_SECOND_STEP) stopped after 40 evaluations: the reference drives NLopt 2.5.0 LD_MMA, the
    product its own on-device restatement of that loop; trajectories within 1e-4 m (north_star)."""
    name, cfg, gm, gold, api = case
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
    X = np.repeat(gold[name + "_x0"][None], 8, axis=0)
    xs, minf = ps.solve_host(gm, X, step=2, maxeval=40, starts_per_problem=8)
    assert np.abs(xs - xs[0]).max() == 0.0  # identical starts -> identical rows
    c_gpu, c_ref = ps.coeff_from_x(xs[0]), gold[name + "_solve40_coeff"]
    T = gold[name + "_seg_time"]
    dev = 0.0
    for s in range(len(T)):  # sample both trajectories every 0.01 s (polynomial_traj.hpp:69-78)
        t = np.arange(0.0, T[s], 0.01)
        pw = t[:, None] ** np.arange(6)[None]
        dev = max(dev, np.abs(pw @ (c_gpu[s].reshape(3, 6) - c_ref[s].reshape(3, 6)).T).max())
    assert dev <= 1e-4, dev
