"""GPU cost+gradient kernel vs the CPU oracle on identical inputs (parity gate: 1e-9 relative).

Reference behaviour under test: GradTrajOptimizer::getCostAndGradient / costFunc
(src/grad_traj_optimizer.cpp:281-448, 554-562 upstream) with the rulings of SURVEY §9.
Parity definition: e_c = |c - c_ref| / |c_ref|, e_g = ||g - g_ref||_inf / ||g_ref||_inf,
max over the batch <= 1e-9 (fp64 SDF)."""
import numpy as np
import pytest

from common import (map_a_batch, map_a_oracle, oracle_map_from_cfg, oracle_problem, rel_cost_err, rel_grad_err)
from grad_traj_optimization_b200 import fixtures as fx

pytestmark = pytest.mark.gpu

TOL = 1e-9  # north_star tolerance for fp64 SDF storage


def _api():
    from grad_traj_optimization_b200 import api
    return api


def _gpu_map_from_oracle(om):
    gm = _api().SdfMap(om.origin, om.resolution, om.size)
    gm.upload_distance(om.distance())
    return gm


def _starts(opt, n_starts, seed):
    lb, ub = opt.bounds()
    return fx.multi_starts(opt.x0(), lb, ub, n_starts, seed=seed)


@pytest.mark.parametrize("name", ["text_input", "opti_node"])
@pytest.mark.parametrize("step", [1, 2])
def test_fixture_dense_path_parity(gtop_lib, name, step):
    """Drop-in path: the oracle's L, R, Df, segment_time go through gtop_problems_create_dense."""
    api = _api()
    cfg = fx.load_cfg(name)
    om = oracle_map_from_cfg(cfg)
    opt = oracle_problem(cfg, om, cfg["waypoints"])
    gm = _gpu_map_from_oracle(om)
    ps = api.ProblemSet.from_dense(opt.seg_time(), opt.L(), opt.R(), opt.Df(), api.make_params(cfg["params"]))
    X = _starts(opt, 64, seed=11)
    c_ref, g_ref, oob_ref, _ = opt.eval_batch(X, step)
    c, g, oob = ps.eval_host(gm, X, step=step, starts_per_problem=X.shape[0])
    assert oob == oob_ref
    assert rel_cost_err(c, c_ref).max() <= TOL
    assert rel_grad_err(g, g_ref).max() <= TOL


@pytest.mark.parametrize("name", ["text_input", "opti_node"])
def test_fixture_native_setup_parity(gtop_lib, name):
    """Native path: waypoints -> gtop_problems_create_from_paths (own O(m) set-up) vs the oracle's
    dense LU set-up; matrices agree to 1e-12 and the objective to 1e-9."""
    api = _api()
    cfg = fx.load_cfg(name)
    om = oracle_map_from_cfg(cfg)
    opt = oracle_problem(cfg, om, cfg["waypoints"])
    gm = _gpu_map_from_oracle(om)
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
    assert (ps.m, ps.n) == (opt.m, opt.n)
    assert np.array_equal(ps.seg_time()[0], opt.seg_time())
    assert np.array_equal(ps.x0()[0], opt.x0())
    lb, ub = opt.bounds()
    assert np.array_equal(ps.bounds()[0][0], lb) and np.array_equal(ps.bounds()[1][0], ub)
    assert np.array_equal(ps.Df(), opt.Df())
    L, R = ps.dense_L(), ps.dense_R()
    assert np.abs(L - opt.L()).max() <= 1e-12 * np.abs(opt.L()).max()
    assert np.abs(R - opt.R()).max() <= 1e-12 * np.abs(opt.R()).max()
    X = _starts(opt, 32, seed=5)
    c_ref, g_ref, _, _ = opt.eval_batch(X, 2)
    c, g, _ = ps.eval_host(gm, X, step=2, starts_per_problem=X.shape[0])
    assert rel_cost_err(c, c_ref).max() <= TOL
    assert rel_grad_err(g, g_ref).max() <= TOL
    assert np.abs(ps.coeff_from_x(X[3]) - opt.coeff_from_x(X[3])).max() <= 1e-11


@pytest.mark.parametrize("m,n_problems,n_starts,ns", [(6, 16, 64, 30),   # config 2 shape (B = 1,024)
                                                      (10, 4, 256, 30),  # config 3 shape, 4 waypoint sets
                                                      (30, 2, 8, 200),   # config 5 shape
                                                      (2, 3, 5, 30), (17, 2, 3, 45), (33, 1, 4, 7),
                                                      # multi-start (DMMA) kernel: smallest m, sample counts that are not
                                                      # multiples of 4, more segments than the ring holds waypoints
                                                      (2, 2, 8, 30), (5, 2, 16, 45), (3, 1, 8, 7), (13, 1, 24, 33)])
def test_map_a_batches_parity(gtop_lib, m, n_problems, n_starts, ns):
    api = _api()
    om, _ = map_a_oracle()
    gm = _gpu_map_from_oracle(om)
    wps, X = map_a_batch(m, n_problems, n_starts)
    prm = dict(fx.PARAMS_OPTI, samples_per_segment=ns)
    ps = api.ProblemSet.from_paths(wps, api.make_params(prm))
    c, g, oob = ps.eval_host(gm, X, step=2, starts_per_problem=n_starts)
    oob_ref = 0
    for p in range(n_problems):
        opt = oracle_problem(prm, om, wps[p])
        sl = slice(p * n_starts, (p + 1) * n_starts)
        c_ref, g_ref, o, _ = opt.eval_batch(X[sl], 2)
        oob_ref += o
        assert rel_cost_err(c[sl], c_ref).max() <= TOL, "problem %d" % p
        assert rel_grad_err(g[sl], g_ref).max() <= TOL, "problem %d" % p
    assert oob == oob_ref


def test_device_path_equals_host_path_and_split_invariance(gtop_lib):
    """Device-resident evaluation, the host (copy-in/out) path and any contiguous split of the batch
    (the multi-GPU partition rule) give bit-identical per-trajectory results."""
    import torch
    api = _api()
    om, _ = map_a_oracle()
    gm = _gpu_map_from_oracle(om)
    wps, X = map_a_batch(10, 3, 40)
    ps = api.ProblemSet.from_paths(wps, api.make_params(fx.PARAMS_OPTI))
    c_h, g_h, _ = ps.eval_host(gm, X, starts_per_problem=40)
    Xd = torch.from_numpy(X).cuda()
    cd = torch.empty(X.shape[0], dtype=torch.float64, device="cuda")
    gd = torch.empty_like(Xd)
    torch.cuda.synchronize()
    ps.eval_device(gm, Xd, cd, gd, starts_per_problem=40, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(cd.cpu().numpy(), c_h) and np.array_equal(gd.cpu().numpy(), g_h)
    for parts in (2, 4, 8):
        bounds = np.linspace(0, X.shape[0], parts + 1).astype(int)
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            c_p, g_p, _ = ps.eval_host(gm, X[lo:hi], traj_offset=int(lo), starts_per_problem=40)
            assert np.array_equal(c_p, c_h[lo:hi]) and np.array_equal(g_p, g_h[lo:hi])


def test_out_of_map_samples_follow_ruling_h3(gtop_lib):
    """Samples outside the map: dist = -1, grad = 0, counted; parity still holds."""
    api = _api()
    cfg = fx.load_cfg("text_input")
    om = oracle_map_from_cfg(cfg)
    opt = oracle_problem(cfg, om, cfg["waypoints"])
    gm = _gpu_map_from_oracle(om)
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
    lb, ub = opt.bounds()
    X = fx.multi_starts(opt.x0(), lb, ub, 16, seed=3, sigma=(3.0, 4.0, 5.0))  # wild starts leave the 5 m tall map
    c_ref, g_ref, oob_ref, _ = opt.eval_batch(X, 2)
    c, g, oob = ps.eval_host(gm, X, starts_per_problem=16)
    assert oob_ref > 0 and oob == oob_ref
    assert rel_cost_err(c, c_ref).max() <= TOL
    assert rel_grad_err(g, g_ref).max() <= TOL


@pytest.mark.parametrize("starts", [4, 8])  # general kernel / multi-start DMMA kernel
def test_small_collision_weight_skips_sample_pass(gtop_lib, starts):
    """|wc| < 1e-4 breaks out of the segment loop (:346)."""
    api = _api()
    cfg = fx.load_cfg("opti_node")
    om = oracle_map_from_cfg(cfg)
    prm = dict(cfg["params"], wc=5e-5)
    opt = oracle_problem(prm, om, cfg["waypoints"])
    gm = _gpu_map_from_oracle(om)
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(prm))
    X = _starts(opt, starts, seed=2)
    c_ref, g_ref, _, _ = opt.eval_batch(X, 2)
    c, g, _ = ps.eval_host(gm, X, starts_per_problem=starts)
    assert rel_cost_err(c, c_ref).max() <= TOL and rel_grad_err(g, g_ref).max() <= TOL


def test_objective_callback_matches_batch(gtop_lib):
    """gtop_objective (NLopt nlopt_func signature; its own low-latency kernel: CTA per trajectory, warp per segment)
    agrees with the rows of the batched evaluation to rounding and meets the parity gate."""
    api = _api()
    cfg = fx.load_cfg("opti_node")
    om = oracle_map_from_cfg(cfg)
    opt = oracle_problem(cfg, om, cfg["waypoints"])
    gm = _gpu_map_from_oracle(om)
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
    X = _starts(opt, 3, seed=8)
    c, g, _ = ps.eval_host(gm, X, starts_per_problem=3)
    obj = api.Objective(gm, ps, 0, step=2)
    for i in range(3):
        ci, gi = obj(X[i])
        assert abs(ci - c[i]) <= 1e-12 * abs(c[i]) and rel_grad_err(gi, g[i]).max() <= 1e-12
    assert obj.evals == 3
    c_ref, g_ref, _ = opt.eval(X[1], 2)
    c1, g1 = obj(X[1])
    assert abs(c1 - c_ref) <= TOL * abs(c_ref) and rel_grad_err(g1, g_ref).max() <= TOL


def test_golden_vectors(gtop_lib):
    """Committed oracle outputs (tests/golden/make_golden.py) — parity without the live oracle."""
    import os
    api = _api()
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "fixtures_eval.npz"))
    for name in ("text_input", "opti_node"):
        cfg = fx.load_cfg(name)
        gm = api.SdfMap(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
        gm.update(cfg["obstacles"])
        ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
        X = gold[name + "_X"]
        for step in (1, 2):
            c, g, _ = ps.eval_host(gm, X, step=step, starts_per_problem=X.shape[0])
            assert rel_cost_err(c, gold["%s_cost_step%d" % (name, step)]).max() <= TOL
            assert rel_grad_err(g, gold["%s_grad_step%d" % (name, step)]).max() <= TOL


def test_both_kernels_agree_on_one_problem(gtop_lib):
    """The same batch through the multi-start DMMA kernel (starts_per_problem % 8 == 0) and through the general
    quarter-warp kernel (any other value; with a single problem the mapping does not change) agrees to rounding
    and both meet the parity gate.  Ragged sizes: B = 37 is neither a multiple of 8 nor of 4."""
    api = _api()
    cfg = fx.load_cfg("opti_node")
    om = oracle_map_from_cfg(cfg)
    opt = oracle_problem(cfg, om, cfg["waypoints"])
    gm = _gpu_map_from_oracle(om)
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
    X = _starts(opt, 37, seed=11)
    for step in (1, 2):
        c_ref, g_ref, _, _ = opt.eval_batch(X, step)
        c_m, g_m, _ = ps.eval_host(gm, X, step=step, starts_per_problem=8)   # k_eval_mma
        c_g, g_g, _ = ps.eval_host(gm, X, step=step, starts_per_problem=37)  # k_eval_traj
        for c, g in ((c_m, g_m), (c_g, g_g)):
            assert rel_cost_err(c, c_ref).max() <= TOL and rel_grad_err(g, g_ref).max() <= TOL
        assert rel_cost_err(c_m, c_g).max() <= 1e-12 and rel_grad_err(g_m, g_g).max() <= 1e-12
    # an offset that is not a multiple of 8 keeps the lane groups aligned on the global index: bit-identical rows
    c_all, g_all, _ = ps.eval_host(gm, X, step=2, starts_per_problem=8)
    c_off, g_off, _ = ps.eval_host(gm, X[5:29], step=2, traj_offset=5, starts_per_problem=8)
    assert np.array_equal(c_off, c_all[5:29]) and np.array_equal(g_off, g_all[5:29])


@pytest.mark.parametrize("starts", [8, 5])  # DMMA kernel / general kernel
def test_map_border_cells_follow_the_clamped_corner_rule(gtop_lib, starts):
    """Samples within half a voxel of the lower map faces have cell index -1 on that axis and the reference
    clamps both corners to voxel 0 (getDistance(Vector3i), sdf_map.cpp:166-174); samples next to the upper
    faces clamp to the last voxel.  The quad-layout gather handles these through its xlow / ylow path."""
    api = _api()
    origin, res, size = np.array([-2.0, -1.5, 0.0]), 0.1, np.array([4.0, 3.0, 2.0])
    from oracle import OracleMap
    om = OracleMap(origin, res, size)
    om.reset()
    rng = fx.Rng(99)
    pts = origin + rng.uniform((400, 3)) * size
    om.set_occupancy(pts[(np.abs(pts[:, 1]) > 0.6) | (pts[:, 2] > 1.2)])  # keep a corridor around y = 0 partly free
    om.update_esdf()
    gm = _gpu_map_from_oracle(om)
    prm = dict(fx.PARAMS_OPTI, bos=0.02, vos=0.05, aos=0.05)  # starts stay next to the straight lines below
    paths = [
        [[-1.9, -1.47, 0.03], [-0.6, -1.48, 0.02], [0.7, -1.46, 0.04], [1.9, -1.47, 0.03]],   # low y, low z faces
        [[-1.97, -1.2, 1.0], [-1.98, -0.2, 1.0], [-1.96, 0.6, 1.0], [-1.97, 1.4, 1.0]],       # low x face
        [[1.96, -1.3, 1.97], [1.97, -0.3, 1.96], [1.98, 0.5, 1.97], [1.97, 1.46, 1.98]],      # upper x, z (and y) faces
    ]
    for wp in paths:
        wp = np.asarray(wp)
        opt = oracle_problem(prm, om, wp)
        ps = api.ProblemSet.from_paths(wp, api.make_params(prm))
        X = _starts(opt, starts, seed=4)
        c_ref, g_ref, oob_ref, _ = opt.eval_batch(X, 2)
        c, g, oob = ps.eval_host(gm, X, step=2, starts_per_problem=starts)
        assert oob == oob_ref
        assert rel_cost_err(c, c_ref).max() <= TOL and rel_grad_err(g, g_ref).max() <= TOL


@pytest.mark.parametrize("starts", [16, 5])
def test_velocity_acceleration_penalty_flag(gtop_lib, starts):
    """the velocity / acceleration penalty block that upstream HEAD has commented out
    (grad_traj_optimizer.cpp:383-407, penalties :517-535) behind gtop_params.vel_acc_penalty (default off).
    Literal restatement on both sides (cv / ca carried from the k = 2 iteration, no sign(v)); step 2 only.
    Parity is unpinned upstream; the gate is the same 1e-9 against the oracle's restatement."""
    api = _api()
    cfg = fx.load_cfg("opti_node")
    om = oracle_map_from_cfg(cfg)
    prm = dict(cfg["params"], alpha_v=2.0, v0=1.5, r_v=0.8, alpha_a=1.5, a0=2.0, r_a=1.2, vel_acc_penalty=1)
    opt = oracle_problem(prm, om, cfg["waypoints"])
    gm = _gpu_map_from_oracle(om)
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(prm))
    X = _starts(opt, starts, seed=21)
    base = oracle_problem(dict(prm, vel_acc_penalty=0), om, cfg["waypoints"])
    for step in (1, 2):
        c_ref, g_ref, _, _ = opt.eval_batch(X, step)
        c, g, _ = ps.eval_host(gm, X, step=step, starts_per_problem=starts)
        assert rel_cost_err(c, c_ref).max() <= TOL and rel_grad_err(g, g_ref).max() <= TOL
        c_off, _, _, _ = base.eval_batch(X, step)
        # the block only acts in OPT_SECOND_STEP, and there it must change the objective noticeably
        assert (np.abs(c_ref - c_off).max() > 1e-3 * np.abs(c_off).max()) == (step == 2)
        # the NLopt callback (single-trajectory kernel) carries the same block
        obj = api.Objective(gm, ps, 0, step=step)
        c1, g1 = obj(X[1])
        assert abs(c1 - c_ref[1]) <= TOL * abs(c_ref[1]) and rel_grad_err(g1, g_ref[1]).max() <= TOL
        obj.close()


def test_full_config3_batch_parity(gtop_lib):
    """BASELINE configs[2] at full size: all 65,536 trajectories (64 waypoint sets x 1,024 starts, m = 10) against the
    oracle (a few seconds on the host cores), plus the 8,192-trajectory shards a rank gets at 8 GPUs: bit-identical rows."""
    api = _api()
    om, _ = map_a_oracle()
    gm = _gpu_map_from_oracle(om)
    wps, X = map_a_batch(10, 64, 1024)
    ps = api.ProblemSet.from_paths(wps, api.make_params(fx.PARAMS_OPTI))
    c, g, oob = ps.eval_host(gm, X, step=2, starts_per_problem=1024)
    assert oob == 0
    ec = eg = 0.0
    for p in range(64):
        opt = oracle_problem(fx.PARAMS_OPTI, om, wps[p])
        sl = slice(1024 * p, 1024 * (p + 1))
        cr, gr, _, _ = opt.eval_batch(X[sl], 2)
        ec, eg = max(ec, rel_cost_err(c[sl], cr).max()), max(eg, rel_grad_err(g[sl], gr).max())
    assert ec <= TOL and eg <= TOL, (ec, eg)
    for r in (0, 3, 7):  # strong-scaling shards: same bits as the full batch
        sl = slice(8192 * r, 8192 * (r + 1))
        cs, gs, _ = ps.eval_host(gm, X[sl], step=2, starts_per_problem=1024, traj_offset=8192 * r)
        assert np.array_equal(cs, c[sl]) and np.array_equal(gs, g[sl])


def test_config5_batch_parity(gtop_lib):
    """BASELINE configs[4] shape: m = 30, 200 samples per segment; 2 waypoint sets x 1,024 starts against the oracle."""
    api = _api()
    om, dist = map_a_oracle()
    gm = _gpu_map_from_oracle(om)
    prm = dict(fx.PARAMS_OPTI, samples_per_segment=200)
    wps = np.stack([fx.random_waypoints(p, 30, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"]) for p in range(2)])
    ps = api.ProblemSet.from_paths(wps, api.make_params(prm))
    x0, (lb, ub) = ps.x0(), ps.bounds()
    X = np.concatenate([fx.multi_starts(x0[p], lb[p], ub[p], 1024, seed=7 + p) for p in range(2)])
    c, g, oob = ps.eval_host(gm, X, step=2, starts_per_problem=1024)
    for p in range(2):
        opt = oracle_problem(prm, om, wps[p])
        sl = slice(1024 * p, 1024 * (p + 1))
        cr, gr, oob_ref, _ = opt.eval_batch(X[sl], 2)
        assert rel_cost_err(c[sl], cr).max() <= TOL and rel_grad_err(g[sl], gr).max() <= TOL


def test_plain_gather_equals_quad_gather(gtop_lib, monkeypatch):
    """Maps too big for the xy-quad copy are gathered straight from the distance field (same corner values, same
    clamping): identical results from every evaluation kernel, including the map border and out-of-map samples."""
    api = _api()
    cfg = fx.load_cfg("text_input")
    om = oracle_map_from_cfg(cfg)
    opt = oracle_problem(cfg, om, cfg["waypoints"])
    X = _starts(opt, 64, seed=21)
    X[5, 2 * opt.num_dp:2 * opt.num_dp + 9:3] += 4.0   # above the ceiling: out-of-map samples
    X[6, 0:9:3] = -9.97                                # along the x = -10 face: border cells (idx = -1)
    res = []
    for limit in ("1024", "0"):
        monkeypatch.setenv("GTOP_QUAD_LIMIT_MB", limit)
        gm = _gpu_map_from_oracle(om)
        assert gm.gather_layout()[0] == (limit != "0")
        ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
        out = [ps.eval_host(gm, X, step=2, starts_per_problem=64),   # k_eval_mma
               ps.eval_host(gm, X[:7], step=2, starts_per_problem=1)]  # k_eval_traj (ragged, 7 trajectories)
        res.append(out)
    for (c0, g0, o0), (c1, g1, o1) in zip(res[0], res[1]):
        assert o0 == o1 and o0 > 0
        assert np.array_equal(c0, c1) and np.array_equal(g0, g1)
    c_ref, g_ref, _, _ = opt.eval_batch(X, 2)
    assert rel_cost_err(res[1][0][0], c_ref).max() <= TOL and rel_grad_err(res[1][0][1], g_ref).max() <= TOL


def test_argument_checks_of_the_evaluation_entry_points(gtop_lib):
    """A negative traj_offset would index problems out of bounds; fields baked into the records at creation cannot be
    changed on a live set; penalty constants outside the kernels' exponent range are refused."""
    from grad_traj_optimization_b200._capi import GtopError
    api = _api()
    cfg = fx.load_cfg("opti_node")
    om = oracle_map_from_cfg(cfg)
    gm = _gpu_map_from_oracle(om)
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
    X = ps.x0()
    with pytest.raises(GtopError, match="traj_offset"):
        ps.eval_host(gm, X, traj_offset=-1)
    with pytest.raises(GtopError, match="fixed at creation"):
        ps.set_params(api.make_params(dict(cfg["params"], bos=1.0)))
    ps.set_params(api.make_params(dict(cfg["params"], wc=2.5)))  # weights may change
    c0, _, _ = ps.eval_host(gm, X)
    ps.set_params(api.make_params(cfg["params"]))
    c1, _, _ = ps.eval_host(gm, X)
    assert c0[0] != c1[0]
    with pytest.raises(GtopError, match="range"):
        api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(dict(cfg["params"], r=1e-5)))
    t = gm.last_update_ms()
    assert len(t) == 6


@pytest.mark.parametrize("vel_acc", [0, 1])
def test_segment_parallel_general_kernel_is_bit_identical(gtop_lib, vel_acc):
    """Small general batches run k_eval_traj_seg (one warp per segment), large ones k_eval_traj (a quarter-warp walks
    the segments): same per-segment code, two-term row sums, cost folded in segment order -> the same bits, so that the
    choice by batch size never shows in the results; both against the oracle."""
    api = _api()
    om, dist = map_a_oracle()
    gm = _gpu_map_from_oracle(om)
    m, nprob, B = 6, 512, 18000
    wps = np.stack([fx.random_waypoints(p, m, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"]) for p in range(nprob)])
    prm = dict(fx.PARAMS_OPTI, vel_acc_penalty=vel_acc, alpha_v=2.0, alpha_a=1.5)
    ps = api.ProblemSet.from_paths(wps, api.make_params(prm), setup="device")
    x0, (lb, ub) = ps.x0(), ps.bounds()
    rng = np.random.default_rng(5)
    pid = np.arange(B) % nprob
    X = np.clip(x0[pid] + rng.normal(0.0, 0.3, (B, ps.n)), lb[pid], ub[pid])
    c, g, _ = ps.eval_host(gm, X, step=2, starts_per_problem=1)        # chunks of 4,096 < 16,384: segment-parallel
    ps.take_out_of_map()
    import torch
    Xd = torch.from_numpy(X).cuda()
    cd, gd = torch.empty(B, dtype=torch.float64, device="cuda"), torch.empty_like(Xd)
    ps.eval_device(gm, Xd, cd, gd, step=2, starts_per_problem=1)      # one launch of 18,000: the sequential kernel
    torch.cuda.synchronize()
    assert np.array_equal(cd.cpu().numpy(), c) and np.array_equal(gd.cpu().numpy(), g)
    for lo, hi in ((0, 1), (5000, 5777), (17990, 18000)):             # ragged slices, offsets that are not multiples of 4
        cs, gs, _ = ps.eval_host(gm, X[lo:hi], step=2, starts_per_problem=1, traj_offset=lo)
        assert np.array_equal(cs, c[lo:hi]) and np.array_equal(gs, g[lo:hi])
    from common import oracle_params
    for p in (0, 7, 511):
        opt = oracle_problem(prm, om, wps[p])
        rows = np.nonzero(pid == p)[0][:4]
        cr, gr, _, _ = opt.eval_batch(X[rows], 2)
        assert rel_cost_err(c[rows], cr).max() <= TOL and rel_grad_err(g[rows], gr).max() <= TOL


def test_big_field_in_hbm_parity(gtop_lib):
    """A map whose field does not fit L2 and whose quad copy would exceed the 1 GiB limit (726 x 718 x 162 voxels = 676 MB:
    the BASELINE configs[3] situation at a size the oracle builds in seconds): the field lives in 8 x 8 x 4 bricks, none of
    the grid dimensions is a multiple of the brick.  GPU EDT bit-exact against the oracle, then distinct problems
    (k_eval_traj_seg) and a multi-start batch (k_eval_mma), both gathering straight from HBM, against the oracle."""
    from oracle import OracleMap
    api = _api()
    res, dims = 0.0625, (726, 718, 162)
    size = np.array(dims, dtype=np.float64) * res
    origin = np.array([-0.5 * size[0], -0.5 * size[1], 0.0])
    gm = api.SdfMap(origin, res, size)
    assert gm.grid == dims and not gm.gather_layout()[0]
    boxes = fx.box_pillars(350, 9, dims, half_width_vox=5, height_vox=(40, 150), margin_vox=40)
    gm.set_occupancy_boxes(boxes), gm.update_esdf()
    occ = np.zeros(dims, dtype=np.int32)
    for b in boxes:
        occ[b[0]:b[3] + 1, b[1]:b[4] + 1, b[2]:b[5] + 1] = 1
    om = OracleMap(origin, res, size)
    assert om.grid == dims
    om.set_occupancy_buffer(occ), om.update_esdf()
    assert np.array_equal(gm.distance(), om.distance())
    # distinct problems, one start each
    m, B = 10, 384
    free = fx.boxes_clearance_fn(boxes, origin, res, 0.5)
    wps = np.stack([fx.random_waypoints_free(p, m, free, 19.0, -18.0 + 34.0 * ((p * 0.6180339887) % 1.0), start_y_span=16.0)
                    for p in range(B)])
    ps = api.ProblemSet.from_paths(wps, api.make_params(fx.PARAMS_OPTI), setup="device")
    x0, (lb, ub) = ps.x0(), ps.bounds()
    X = np.clip(x0 + np.random.default_rng(3).normal(0.0, 0.25, x0.shape), lb, ub)
    c, g, oob = ps.eval_host(gm, X, step=2, starts_per_problem=1)
    assert oob == 0
    for p in range(0, B, 7):
        cr, gr, _, _ = oracle_problem(fx.PARAMS_OPTI, om, wps[p]).eval_batch(X[p:p + 1], 2)
        assert rel_cost_err(c[p:p + 1], cr).max() <= TOL and rel_grad_err(g[p:p + 1], gr).max() <= TOL
    # multi-start batch of two of the problems
    ps2 = api.ProblemSet.from_paths(wps[:2], api.make_params(fx.PARAMS_OPTI))
    Xm = np.concatenate([fx.multi_starts(x0[p], lb[p], ub[p], 256, seed=11 + p) for p in range(2)])
    cm, gm_, _ = ps2.eval_host(gm, Xm, step=2, starts_per_problem=256)
    for p in range(2):
        sl = slice(256 * p, 256 * (p + 1))
        cr, gr, _, _ = oracle_problem(fx.PARAMS_OPTI, om, wps[p]).eval_batch(Xm[sl], 2)
        assert rel_cost_err(cm[sl], cr).max() <= TOL and rel_grad_err(gm_[sl], gr).max() <= TOL


def test_pinned_host_buffers_are_evaluated_in_place(gtop_lib, monkeypatch):
    """gtop_eval_batch_host on pinned buffers lets the multi-start kernel read x and write cost / gradient across PCIe
    directly (no staging copies); pageable buffers and GTOP_HOST_ZEROCOPY=0 take the chunked staging pipeline: same bits,
    same out-of-map count, ragged offsets included."""
    import torch
    api = _api()
    cfg = fx.load_cfg("opti_node")
    om = oracle_map_from_cfg(cfg)
    gm = _gpu_map_from_oracle(om)
    opt = oracle_problem(cfg, om, cfg["waypoints"])
    X = _starts(opt, 4096, seed=5)
    X[7, 2 * opt.num_dp:2 * opt.num_dp + 9:3] += 4.0   # above the ceiling: out-of-map samples
    ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
    c0, g0, o0 = ps.eval_host(gm, X, step=2, starts_per_problem=4096)          # pageable numpy buffers
    assert o0 > 0
    Xh = torch.from_numpy(X).pin_memory()
    Gh, Ch = torch.empty(X.shape, dtype=torch.float64).pin_memory(), torch.empty((X.shape[0],), dtype=torch.float64).pin_memory()
    n0 = gtop_lib.gtop_kernel_launch_count()
    _, _, o1 = ps.eval_host(gm, Xh, step=2, starts_per_problem=4096, cost=Ch, grad=Gh)
    assert gtop_lib.gtop_kernel_launch_count() - n0 == 1                       # one launch, no chunks
    assert o1 == o0 and np.array_equal(Ch.numpy(), c0) and np.array_equal(Gh.numpy(), g0)
    Gh.zero_(), Ch.zero_()
    _, _, o2 = ps.eval_host(gm, Xh[1000:3001], step=2, starts_per_problem=4096, traj_offset=1000, cost=Ch[1000:3001], grad=Gh[1000:3001])
    assert np.array_equal(Ch.numpy()[1000:3001], c0[1000:3001]) and np.array_equal(Gh.numpy()[1000:3001], g0[1000:3001])
    assert not Ch.numpy()[:1000].any() and not Gh.numpy()[3001:].any()
    c_ref, g_ref, _, _ = opt.eval_batch(X[:64], 2)
    assert rel_cost_err(c0[:64], c_ref).max() <= TOL and rel_grad_err(g0[:64], g_ref).max() <= TOL
