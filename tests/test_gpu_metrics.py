"""Post-hoc trajectory metrics on the device against the oracle's PolynomialTraj restatement.

Reference behaviour under test: what opti_node computes from an optimised trajectory (src/opti_node.cpp:104-144)
through include/grad_traj_optimization/polynomial_traj.hpp — getTraj/getLength (:70-91), getJerk (:109-139),
getMeanAndMaxVel/Acc (:141-203, literal pow(ts, i)), getAccCost (:95-107).  Floating-point quantities: gate 1e-9
relative (measured ~1e-13); the number of trajectory points (the 0.01 s clock) is reproduced exactly, which the length
gate would expose (one point more or less changes the length by ~1e-3 relative)."""
import os

import numpy as np
import pytest

from common import map_a_batch, map_a_oracle, oracle_map_from_cfg, oracle_problem
from grad_traj_optimization_b200 import fixtures as fx
from oracle import traj_metrics

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _api():
    from grad_traj_optimization_b200 import api
    return api


def _check(met, ref, what):
    """relative gate; the velocity / acceleration entries can be pure cancellation noise (the straight-line start has
    zero velocity and acceleration at every waypoint, which is where the reference's literal pow(ts, i) samples them), so
    they are gated against the trajectory's own speed scale instead of their value"""
    speed = ref[1] / ref[0]  # length / time_sum
    floor = np.array([0.0, 0.0, 0.0, speed, speed, speed / 0.1, speed / 0.1, 0.0])
    err = np.abs(met - ref) / np.maximum(np.maximum(np.abs(ref), floor), 1e-300)
    assert err.max() <= TOL, "%s: metric %d off by %.3g (gpu %r, oracle %r)" % (
        what, int(err.argmax()), err.max(), met[err.argmax()], ref[err.argmax()])


@pytest.mark.parametrize("name", ["text_input", "opti_node"])
def test_fixture_metrics(gtop_lib, name):
    api = _api()
    cfg = fx.load_cfg(name)
    om = oracle_map_from_cfg(cfg)
    opt = oracle_problem(cfg, om, cfg["waypoints"])
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
    lb, ub = opt.bounds()
    X = fx.multi_starts(opt.x0(), lb, ub, 9, seed=5)
    met = ps.traj_metrics_host(X, starts_per_problem=9)
    assert met.shape == (9, 8)
    for b in range(9):
        _check(met[b], traj_metrics(opt.coeff_from_x(X[b]), opt.seg_time()), "%s start %d" % (name, b))
    # committed golden vectors (tests/golden/make_golden.py): parity without the live oracle
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "fixtures_eval.npz"))
    Xg = gold[name + "_X"]
    mg = ps.traj_metrics_host(Xg, starts_per_problem=Xg.shape[0])
    for b in range(Xg.shape[0]):
        _check(mg[b], gold[name + "_metrics"][b], "%s golden start %d" % (name, b))


@pytest.mark.parametrize("m,n_problems,n_starts,setup", [(6, 8, 4, "host"), (10, 4, 8, "device"), (30, 2, 3, "host"),
                                                         (2, 3, 2, "device"), (40, 1, 5, "host")])
def test_map_a_metrics(gtop_lib, m, n_problems, n_starts, setup):
    api = _api()
    om, _ = map_a_oracle()
    wps, X = map_a_batch(m, n_problems, n_starts)
    ps = api.ProblemSet.from_paths(wps, api.make_params(dict(fx.PARAMS_OPTI)), setup=setup)
    met = ps.traj_metrics_host(X, starts_per_problem=n_starts)
    for p in range(n_problems):
        opt = oracle_problem(dict(fx.PARAMS_OPTI), om, wps[p])
        for k in range(n_starts):
            b = p * n_starts + k
            _check(met[b], traj_metrics(opt.coeff_from_x(X[b]), opt.seg_time()), "problem %d start %d" % (p, k))


def test_metrics_device_path_and_offsets(gtop_lib):
    """device-resident call, traj_offset mapping, and ranking of solved multi-starts without leaving the GPU"""
    import torch
    api = _api()
    om, _ = map_a_oracle()
    gm = api.SdfMap(om.origin, om.resolution, om.size)
    gm.upload_distance(om.distance())
    wps, X = map_a_batch(6, 4, 16)
    ps = api.ProblemSet.from_paths(wps, api.make_params(dict(fx.PARAMS_OPTI)))
    full = ps.traj_metrics_host(X, starts_per_problem=16)
    Xd = torch.from_numpy(X[24:56].copy()).cuda()
    out = torch.empty(32, 8, dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    ps.traj_metrics_device(Xd, out, traj_offset=24, starts_per_problem=16, stream=st.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), full[24:56])
    # after a solve the best start of each problem by (cost, then length) can be picked on the device
    Xs, minf = ps.solve_host(gm, X, step=2, maxeval=20, starts_per_problem=16)
    met = ps.traj_metrics_host(Xs, starts_per_problem=16)
    assert np.all(met[:, 1] > 0) and np.all(met[:, 4] >= met[:, 3] - 1e-12) and np.all(np.isfinite(met))
    t_sum = ps.seg_time().sum(axis=1)
    assert np.allclose(met[:, 0].reshape(4, 16), t_sum[:, None], rtol=1e-15)
