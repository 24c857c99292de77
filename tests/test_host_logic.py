"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol, argument
validation fails loudly without a device, fixtures are deterministic, sharding (gloo, world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from common import map_a_batch, map_a_oracle, oracle_problem
from grad_traj_optimization_b200 import _capi, fixtures as fx, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_declared_in_header():
    hdr = open(os.path.join(ROOT, "include", "gtop_c.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gtop_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    L = _capi.lib()  # raises AttributeError if a declared symbol is not exported
    raw = C.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name)
    assert L.gtop_abi_version() == 1


def test_params_struct_matches_header_and_defaults():
    p = _capi.GtopParams()
    _capi.lib().gtop_params_default(C.byref(p))
    assert C.sizeof(p) == 20 * 8 + 4 * 4
    assert (p.algorithm, p.samples_per_segment, p.vel_acc_penalty) == (24, 30, 0)
    assert p.bos == 3.0 and p.vos == 8.0 and p.aos == 10.0


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point must fail with an error code, never compute on the CPU."""
    L = _capi.lib()
    if L.gtop_device_count() > 0:
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    o = (C.c_double * 3)(0, 0, 0)
    s = (C.c_double * 3)(1, 1, 1)
    rc = L.gtop_sdf_create(o, 0.1, s, 0, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in L.gtop_last_error()
    p = _capi.GtopParams()
    L.gtop_params_default(C.byref(p))
    w = np.zeros((3, 3))
    w[:, 0] = [0, 1, 2]
    rc = L.gtop_problems_create_from_paths(1, 2, w.ctypes.data_as(C.POINTER(C.c_double)), C.byref(p), 0, C.byref(h))
    assert rc != 0


def test_new_entry_points_reject_bad_arguments_and_never_fall_back():
    """Argument checks of the entry points added for SURVEY §8f run before any device work; without a device the
    computing ones fail with an error code (no CPU fallback)."""
    L = _capi.lib()
    dp = C.POINTER(C.c_double)
    z3 = (C.c_double * 3)(0, 0, 0)
    assert L.gtop_traj_metrics_host(None, z3, 1, 0, 1, z3) < 0 and b"null" in L.gtop_last_error()
    assert L.gtop_sdf_query_dynamic(None, None, 0, 0.0, z3, 1, z3, z3, None) < 0
    assert L.gtop_sdf_set_update_range(None, z3, z3) < 0 and L.gtop_sdf_reset_range(None, z3, z3) < 0
    lay = (C.c_int32 * 16)()
    p = _capi.GtopParams()
    L.gtop_params_default(C.byref(p))
    w = np.zeros((3, 3))
    w[:, 0] = [0, 1, 2]
    assert L.gtop_debug_pack_record(1, w.ctypes.data_as(dp), C.byref(p), None, 0, lay) < 0           # m >= 2
    assert L.gtop_debug_pack_record(2, w.ctypes.data_as(dp), C.byref(p), None, 0, lay) == 0
    rec = np.zeros(8)
    assert L.gtop_debug_pack_record(2, w.ctypes.data_as(dp), C.byref(p), rec.ctypes.data_as(dp), 8, lay) < 0  # buffer too small
    bad = np.zeros((3, 3))  # coincident waypoints: zero segment time
    rec = np.zeros(lay[5])
    assert L.gtop_debug_pack_record(2, bad.ctypes.data_as(dp), C.byref(p), rec.ctypes.data_as(dp), lay[5], lay) < 0
    assert b"segment time" in L.gtop_last_error()
    if L.gtop_device_count() == 0:
        h = C.c_void_p()
        rc = L.gtop_problems_create_from_paths_device(1, 2, w.ctypes.data_as(dp), C.byref(p), 0, C.byref(h))
        assert rc != 0 and b"no CPU fallback" in L.gtop_last_error()


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under grad_traj_optimization_b200/ or include/ may reference it."""
    bad = []
    for base in ("grad_traj_optimization_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle\b|oracle/|gtop_oracle|libgtop_oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_fixture_configs_expand_like_the_demos():
    ti, on = fx.load_cfg("text_input"), fx.load_cfg("opti_node")
    assert ti["obstacles"].shape == (36 * 9 * 25, 3) and ti["waypoints"].shape == (8, 3)
    assert on["obstacles"].shape == (3100, 3) and on["waypoints"].shape == (11, 3)
    assert ti["params"]["ws"] == 200.0 and on["params"]["wc"] == 5.0
    zs = np.unique(ti["obstacles"][:, 2])
    assert len(zs) == 36 and zs[0] == 0.0 and zs[-1] < 3.5  # float z accumulation (H13)


def test_rng_and_workloads_are_deterministic():
    a, b = fx.Rng(42).uniform((5,)), fx.Rng(42).uniform((5,))
    assert np.array_equal(a, b) and np.all((a >= 0) & (a < 1))
    assert abs(fx.Rng(1).normal((20000,)).std() - 1.0) < 0.02
    c1, c2 = fx.map_a_pillar_centres(), fx.map_a_pillar_centres()
    assert np.array_equal(c1, c2) and c1.shape == (120, 2)
    d = np.linalg.norm(c1[:, None] - c1[None], axis=2) + np.eye(120) * 10
    assert d.min() > 1.2
    _, dist = map_a_oracle()
    w1 = fx.random_waypoints(3, 10, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"])
    w2 = fx.random_waypoints(3, 10, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"])
    assert np.array_equal(w1, w2) and w1.shape == (11, 3)
    assert np.all(np.abs(w1[:, :2]) <= 8.5) and np.all(np.abs(w1[:, 2] - 2.0) <= 0.5)
    steps = np.linalg.norm(np.diff(w1[:, :2], axis=0), axis=1)
    assert np.all((steps >= 1.2 - 1e-9) & (steps <= 2.5 + 1e-9))
    x0, lb, ub = fx.straight_line_x0(w1)
    X = fx.multi_starts(x0, lb, ub, 16, seed=7)
    assert np.array_equal(X[0], x0) and np.all(X >= lb) and np.all(X <= ub)


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = sharding.shard_range(total, world, r)
                assert 0 <= hi - lo <= total // world + 1
                cover += list(range(lo, hi)) if total < 100 else [lo, hi]
            if total < 100:
                assert cover == list(range(total))
            else:
                assert cover[0] == 0 and cover[-1] == total and all(cover[i] == cover[i + 1] for i in range(1, len(cover) - 1, 2))
    assert sharding.problem_of(65536 * 3 + 2047, 1024, 64) == 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from common import map_a_batch, map_a_oracle, oracle_problem
from grad_traj_optimization_b200 import fixtures as fx, sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_problems, n_starts, m = 2, 6, 4
om, _ = map_a_oracle()
wps, X = map_a_batch(m, n_problems, n_starts)
lo, hi = sharding.shard_range(X.shape[0], world, rank)
cost = np.zeros(X.shape[0])
for b in range(lo, hi):            # rank-local block, global problem mapping via the offset
    p = sharding.problem_of(b, n_starts, n_problems)
    cost[b] = oracle_problem(fx.PARAMS_OPTI, om, wps[p]).eval(X[b], 2)[0]
t = torch.from_numpy(cost)
dist.all_reduce(t)                 # host-side gather of the disjoint blocks
tmax = sharding.max_over_ranks(1.0 + rank, dist)
if rank == 0:
    full = np.concatenate([oracle_problem(fx.PARAMS_OPTI, om, wps[p]).eval_batch(X[p*n_starts:(p+1)*n_starts], 2)[0] for p in range(n_problems)])
    assert np.array_equal(t.numpy(), full), "sharded result differs from the unsharded one"
    assert tmax == float(world)
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding_matches_unsharded(tmp_path):
    """N>1 host logic on CPU: 2 ranks (gloo), contiguous blocks, offset-based problem mapping, MAX timing reduce."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29513", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("m,ns", [(10, 30), (7, 30), (4, 29), (5, 32), (2, 7), (3, 45)])
def test_pair_schedule_tables_of_the_record(gtop_lib_cpu, m, ns):
    """The multi-start kernel walks segments in pairs with consecutively numbered samples; the record's TQ / PL tables
    carry that schedule (csrc/gtop_types.cuh).  Host-only check of their structure against the plain sample table:
    slot q = 4u + j of pair p holds sample q of segment 2p (q < ns_a) or q - ns_a of segment 2p+1; TQ = (dt, t, t^2..t^5),
    idle slots zero; PL = dt * (phi_r(t), phi_r'(t)) for the slot's segment, the second segment's lanes of the mixed
    iteration in the pair's extra record, zero for the other segment's lanes and for columns 6, 7."""
    import ctypes as C
    from grad_traj_optimization_b200 import api, fixtures as fx
    L = gtop_lib_cpu
    prm = api.make_params(dict(fx.PARAMS_OPTI, samples_per_segment=ns))
    rng = fx.Rng(3 * m + ns)
    wp = np.cumsum(np.concatenate([[[-8.0, 0.0, 2.0]], 0.8 + rng.uniform((m, 3))]), axis=0)
    lay = (C.c_int32 * 16)()
    dp = C.POINTER(C.c_double)
    assert L.gtop_debug_pack_record(m, wp.ctypes.data_as(dp), C.byref(prm), None, 0, lay) == 0
    (m_, ns_cap, nsp, npair, nit2, stride, o_dt, o_ns, o_tt, o_tq, o_pl, o_litp, o_linv, o_T, o_rc, o_rg) = list(lay)
    assert (m_, ns_cap, npair, nit2) == (m, ns, (m + 1) // 2, (2 * ns + 3) // 4)
    rec = np.zeros(stride)
    assert L.gtop_debug_pack_record(m, wp.ctypes.data_as(dp), C.byref(prm), rec.ctypes.data_as(dp), stride, lay) == 0
    dt, nsv = rec[o_dt:o_dt + m], rec[o_ns:o_ns + m].astype(int)
    tt = rec[o_tt:o_tt + m * 5 * ns].reshape(m, 5, ns)
    Linv = rec[o_linv:o_linv + 36 * m].reshape(m, 6, 6)
    TQ = rec[o_tq:o_tq + npair * nit2 * 24].reshape(npair, nit2, 3, 4, 2)
    PLt = rec[o_pl:o_pl + npair * (nit2 + 1) * 64].reshape(npair, nit2 + 1, 8, 4, 2)   # [.., g, j, (A,B)]
    T = rec[o_T:o_T + m]
    assert np.allclose(dt * ns, T, rtol=1e-15) and np.all(nsv >= 1) and np.all(nsv <= ns)
    seen_tq = np.zeros(TQ.shape[:2] + (4,), bool)
    exp_pl = np.zeros_like(PLt)
    for s in range(m):
        p, second = s // 2, s & 1
        ns_first = nsv[s - 1] if second else nsv[s]
        for k in range(nsv[s]):
            q = ns_first + k if second else k
            u, j = q // 4, q % 4
            t = tt[s, 0, k]
            pw = np.array([1.0, t, tt[s, 1, k], tt[s, 2, k], tt[s, 3, k], tt[s, 4, k]])
            assert np.array_equal(TQ[p, u, :, j, :].ravel(), np.concatenate([[dt[s]], pw[1:]]))
            seen_tq[p, u, j] = True
            mixed_iter = (ns_first % 4 != 0) and u == ns_first // 4
            recpl = nit2 if (second and mixed_iter) else u
            for g in range(6):
                r = (g + 3) % 6 if (s & 1) else g
                a = sum(pw[i] * Linv[s, i, r] for i in range(6))
                b = sum(i * pw[i - 1] * Linv[s, i, r] for i in range(1, 6))
                exp_pl[p, recpl, g, j] = (dt[s] * a, dt[s] * b)
    assert np.all(np.moveaxis(TQ, 2, 3)[~seen_tq] == 0.0)   # idle slots
    assert np.allclose(PLt, exp_pl, rtol=1e-13, atol=1e-300)
    assert np.all(PLt[exp_pl == 0.0] == 0.0)                # zeros exactly where nothing is scheduled
