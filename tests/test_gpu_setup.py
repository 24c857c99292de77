"""Set-up of problem sets on the device against the host set-up and the oracle.

Reference behaviour under test: GradTrajOptimizer::setPath / setKinoPath (src/grad_traj_optimizer.cpp:35-110) with
TrajectoryGenerator's per-segment matrices (src/qp_generator.cpp:181-236, :357-451).  The device path runs the same
source as the host path (csrc/setup_core.cuh) without FMA contraction; only pow() differs (correctly rounded on the
device, libm on the host: 1 ulp apart in ~0.08 % of the calls), so the records agree to rounding level.  Gate of §8f-2: <= 1e-12 against the LU-based set-up; evaluations keep the 1e-9 gate against the oracle."""
import numpy as np
import pytest

from common import map_a_batch, map_a_oracle, oracle_problem, rel_cost_err, rel_grad_err
from grad_traj_optimization_b200 import fixtures as fx

pytestmark = pytest.mark.gpu


def _api():
    from grad_traj_optimization_b200 import api
    return api


def _gpu_map_from_oracle(om):
    gm = _api().SdfMap(om.origin, om.resolution, om.size)
    gm.upload_distance(om.distance())
    return gm


def _ulps(a, b):
    """distance in units of the last place between two float64 arrays of the same sign structure"""
    ia, ib = a.view(np.int64), b.view(np.int64)
    return np.abs(ia - ib)


@pytest.mark.parametrize("m,n_problems,ns", [(6, 64, 30), (10, 32, 30), (30, 4, 200), (2, 5, 7), (17, 3, 45)])
def test_device_records_equal_host_records(gtop_lib, m, n_problems, ns):
    api = _api()
    wps, _ = map_a_batch(m, n_problems, 1)
    prm = api.make_params(dict(fx.PARAMS_OPTI, samples_per_segment=ns))
    ph = api.ProblemSet.from_paths(wps, prm)
    pd = api.ProblemSet.from_paths(wps, prm, setup="device")
    assert np.array_equal(ph.x0(), pd.x0()) and np.array_equal(ph.seg_time(), pd.seg_time())
    for a, b in zip(ph.bounds(), pd.bounds()):
        assert np.array_equal(a, b)
    off = ph.layout_offsets()
    n_pow_diff = n_pow = 0
    for p in range(n_problems):
        rh, rd = ph.record(p), pd.record(p)
        assert rh.shape == rd.shape
        # libm's pow differs from the correctly rounded power by 1 ulp in ~0.08 % of the calls; what is computed from
        # such a power (blocks, bands, fragment tables) moves at rounding level
        assert (np.abs(rh - rd) <= 1e-12 * np.abs(rh) + 1e-14).all()
        tt_h, tt_d = rh[off["tt"]:off["Rc"]], rd[off["tt"]:off["Rc"]]
        assert _ulps(tt_h, tt_d).max() <= 1
        n_pow_diff += int((tt_h != tt_d).sum())
        n_pow += tt_h.size
    assert n_pow_diff <= 0.01 * n_pow, "%d of %d powers differ" % (n_pow_diff, n_pow)
    # the accessors that need the blocks on the host set the problem up on first use
    assert np.array_equal(ph.dense_L(0), pd.dense_L(0)) and np.array_equal(ph.dense_R(n_problems - 1), pd.dense_R(n_problems - 1))


@pytest.mark.parametrize("m,n_problems,n_starts,ns", [(6, 16, 1, 30), (10, 4, 64, 30), (30, 2, 8, 200), (5, 3, 5, 33)])
def test_device_setup_evaluations_match_oracle_and_host_setup(gtop_lib, m, n_problems, n_starts, ns):
    api = _api()
    om, _ = map_a_oracle()
    gm = _gpu_map_from_oracle(om)
    wps, X = map_a_batch(m, n_problems, n_starts)
    prm = dict(fx.PARAMS_OPTI, samples_per_segment=ns)
    pd = api.ProblemSet.from_paths(wps, api.make_params(prm), setup="device")
    ph = api.ProblemSet.from_paths(wps, api.make_params(prm))
    c, g, oob = pd.eval_host(gm, X, step=2, starts_per_problem=n_starts)
    ch, gh, _ = ph.eval_host(gm, X, step=2, starts_per_problem=n_starts)
    assert rel_cost_err(c, ch).max() <= 1e-12 and rel_grad_err(g, gh).max() <= 1e-12
    for p in range(n_problems):
        opt = oracle_problem(prm, om, wps[p])
        sl = slice(p * n_starts, (p + 1) * n_starts)
        c_ref, g_ref, _, _ = opt.eval_batch(X[sl], 2)
        assert rel_cost_err(c[sl], c_ref).max() <= 1e-9 and rel_grad_err(g[sl], g_ref).max() <= 1e-9


def test_device_setup_kino_and_errors(gtop_lib):
    from grad_traj_optimization_b200._capi import GtopError
    api = _api()
    m, n_problems = 7, 6
    wps, _ = map_a_batch(m, n_problems, 1)
    rng = fx.Rng(77)
    vel = 0.5 * (rng.uniform((n_problems, m + 1, 3)) - 0.5)
    acc = 0.3 * (rng.uniform((n_problems, m + 1, 3)) - 0.5)
    time = 0.6 + rng.uniform((n_problems, m))
    prm = api.make_params(dict(fx.PARAMS_OPTI))
    ph = api.ProblemSet.from_kino(wps, vel, acc, time, prm)
    pd = api.ProblemSet.from_kino(wps, vel, acc, time, prm, setup="device")
    assert np.array_equal(ph.x0(), pd.x0()) and np.array_equal(ph.seg_time(), pd.seg_time())
    for p in range(n_problems):
        rh, rd = ph.record(p), pd.record(p)
        assert (np.abs(rh - rd) <= 1e-12 * np.abs(rh) + 1e-14).all()
    bad = wps.copy()
    bad[3, 4] = bad[3, 3]  # coincident waypoints: zero segment time (the reference would divide by zero in A^-1)
    with pytest.raises(GtopError, match="problem 3"):
        api.ProblemSet.from_paths(bad, prm, setup="device")


def test_device_setup_solve_matches_host_setup(gtop_lib):
    """the batched optimiser loop takes its NLopt box from the problem set: same solves from either set-up"""
    api = _api()
    om, _ = map_a_oracle()
    gm = _gpu_map_from_oracle(om)
    wps, X = map_a_batch(6, 8, 8)
    prm = api.make_params(dict(fx.PARAMS_OPTI))
    out = []
    for setup in ("host", "device"):
        ps = api.ProblemSet.from_paths(wps, prm, setup=setup)
        out.append(ps.solve_host(gm, X.copy(), step=2, maxeval=30, starts_per_problem=8))
    (xh, fh), (xd, fd) = out
    assert np.abs(xh - xd).max() <= 1e-6 and np.abs(fh - fd).max() <= 1e-6 * np.abs(fh).max()
