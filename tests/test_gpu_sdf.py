"""GPU voxel-map kernels vs the CPU oracle: occupancy scatter, exact EDT (bit-exact), trilinear query.

Reference behaviour under test: SDFMap::setOccupancy / updateESDF3d / getDistWithGradTrilinear
(src/sdf_map.cpp:80-99, 310-368, 185-242 upstream)."""
import numpy as np
import pytest

from common import oracle_map_from_cfg
from grad_traj_optimization_b200 import fixtures as fx
from oracle import OracleMap

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["quad", "bricks"])
def field_layout(request, monkeypatch):
    """Every map test runs on both layouts of the distance field: linear + xy-quad copy (maps under GTOP_QUAD_LIMIT_MB)
    and 8 x 8 x 4 bricks without a copy (big maps; forced here with a limit of 0)."""
    monkeypatch.setenv("GTOP_QUAD_LIMIT_MB", "1024" if request.param == "quad" else "0")
    return request.param


def _gpu_map(origin, res, size):
    import os
    from grad_traj_optimization_b200.api import SdfMap
    gm = SdfMap(origin, res, size)
    assert gm.gather_layout()[0] == (os.environ["GTOP_QUAD_LIMIT_MB"] != "0")
    return gm


def _both(origin, res, size, pts):
    om = OracleMap(origin, res, size)
    om.reset(), om.set_occupancy(pts), om.update_esdf()
    gm = _gpu_map(origin, res, size)
    gm.update(pts)
    return om, gm


@pytest.mark.parametrize("name", ["text_input", "opti_node"])
def test_edt_fixture_maps_bit_exact(gtop_lib, name):
    cfg = fx.load_cfg(name)
    om, gm = _both(cfg["map_origin"], cfg["resolution"], cfg["map_size"], cfg["obstacles"])
    assert gm.grid == om.grid
    assert np.array_equal(gm.occupancy(), om.occupancy())
    d_gpu, d_ref = gm.distance(), om.distance()
    assert np.array_equal(d_gpu, d_ref), "EDT differs: max |diff| = %g" % np.abs(d_gpu - d_ref).max()


def test_edt_map_a_bit_exact(gtop_lib):
    om, gm = _both(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"], fx.map_a_obstacles())
    assert np.array_equal(gm.occupancy(), om.occupancy())
    assert np.array_equal(gm.distance(), om.distance())
    ms = gm.last_edt_ms()
    assert ms[3] > 0.0


@pytest.mark.parametrize("dims,nocc,seed", [((7, 5, 3), 3, 1), ((33, 17, 40), 25, 2), ((1, 1, 9), 2, 3), ((40, 1, 1), 3, 4),
                                           ((64, 64, 33), 200, 5), ((20, 70, 65), 1, 6), ((300, 9, 5), 7, 7),
                                           # z-pass variants: ballot words in registers up to nz = 128, two-sweep kernel above
                                           ((4, 3, 130), 5, 8), ((3, 3, 300), 6, 9), ((2, 3, 530), 4, 10),
                                           # vectorised z pass (nz % 4 == 0, 128 < nz <= 512): 2 and 4 chunks, edge sizes
                                           ((5, 2, 256), 9, 11), ((2, 2, 512), 3, 12), ((3, 2, 132), 2, 13), ((2, 2, 260), 1, 14)])
def test_edt_random_small_grids_vs_bruteforce(gtop_lib, dims, nocc, seed):
    res = 0.25
    size = np.array(dims, dtype=np.float64) * res
    origin = np.array([-1.0, 2.0, 0.5])
    rng = fx.Rng(seed)
    idx = np.stack([np.floor(rng.uniform((nocc,)) * d).astype(int) for d in dims], axis=1)
    occ = np.zeros(dims, dtype=np.int32)
    occ[idx[:, 0], idx[:, 1], idx[:, 2]] = 1
    om = OracleMap(origin, res, size)
    assert om.grid == tuple(dims)
    om.set_occupancy_buffer(occ)
    om.update_esdf()
    gm = _gpu_map(origin, res, size)
    gm.upload_occupancy(occ)
    gm.update_esdf()
    d_gpu = gm.distance()
    assert np.array_equal(d_gpu, om.distance())
    assert np.array_equal(d_gpu, om.brute_edt())


@pytest.mark.parametrize("dims,nocc,seed,windowed", [
    # rows longer than every template bucket of the y / x passes of csrc/sdf_kernels.cu: <= 2048 (the 2000-voxel rows of
    # BASELINE config 4), <= 4096, <= 16384 (64-bit intersection arithmetic above 8192), on both row axes
    ((2000, 8, 8), 40, 21, False), ((8, 2000, 8), 40, 22, False), ((4000, 4, 4), 30, 23, False), ((4, 4000, 4), 30, 24, False),
    ((9000, 2, 2), 25, 25, False), ((2, 9000, 2), 25, 26, False), ((2000, 8, 8), 40, 27, True), ((4, 4000, 4), 30, 28, True),
    ((9000, 2, 2), 25, 29, True), ((2000, 3, 5), 1, 30, False), ((2000, 2, 2), 6000, 31, False),
    # a column taller than 8192 voxels: squared distances beyond the fast integer square root's range
    ((2, 3, 16000), 1, 32, False)])
def test_edt_long_rows_vs_oracle_and_bruteforce(gtop_lib, dims, nocc, seed, windowed):
    """sdf_map.cpp:266-368 on grids whose x / y rows are as long as the large map's (and longer)."""
    res = 0.25  # (dims * res is exact in binary: ceil(size / res) gives back dims)
    size = np.array(dims, dtype=np.float64) * res
    origin = np.array([-3.0, 1.0, 0.0])
    rng = fx.Rng(seed)
    idx = np.stack([np.floor(rng.uniform((nocc,)) * d).astype(int) for d in dims], axis=1)
    occ = np.zeros(dims, dtype=np.int32)
    occ[idx[:, 0], idx[:, 1], idx[:, 2]] = 1
    om = OracleMap(origin, res, size)
    assert om.grid == tuple(dims)
    gm = _gpu_map(origin, res, size)
    om.set_occupancy_buffer(occ), gm.upload_occupancy(occ)
    if windowed:  # setUpdateRange over most of the long axis, then a full rebuild on top (min with the old field)
        lo, hi = origin + 0.1 * size, origin + 0.93 * size
        for mp in (om, gm):
            mp.set_update_range(lo, hi), mp.update_esdf()
        assert np.array_equal(gm.distance(), om.distance())
    for mp in (om, gm):
        mp.update_esdf()
    d_gpu = gm.distance()
    assert np.array_equal(d_gpu, om.distance())
    assert np.array_equal(d_gpu, om.brute_edt())


def test_distance_device_pointer_only_for_the_linear_layout(gtop_lib, field_layout):
    """gtop_sdf_distance_device_ptr hands out the buffer in the reference's index order; a bricked field refuses (its
    callers would read the wrong voxels) and points at the download."""
    import ctypes as C
    gm = _gpu_map([0, 0, 0], 0.5, [8, 6, 4])
    p = C.c_void_p()
    rc = gtop_lib.gtop_sdf_distance_device_ptr(gm.h, C.byref(p))
    if field_layout == "quad":
        assert rc == 0 and p.value
    else:
        assert rc != 0 and b"bricks" in gtop_lib.gtop_last_error()


def test_edt_empty_map_is_10000_everywhere(gtop_lib):
    gm = _gpu_map([0, 0, 0], 0.5, [8, 6, 4])
    gm.reset()
    gm.update_esdf()
    assert np.all(gm.distance() == 10000.0)  # sdf_map.cpp:22,358-360


def test_edt_distances_capped_by_the_reset_value(gtop_lib):
    """min(resolution*sqrt(val), old) with old = 10000 after resetBuffer (sdf_map.cpp:22, :358-360): on a grid coarse enough
    for distances beyond 10000 the far voxels keep the reset value (the x pass only drops the min when no distance of the
    grid can reach it)."""
    origin, res, size = [0.0, 0.0, 0.0], 1500.0, [9 * 1500.0, 7 * 1500.0, 5 * 1500.0]
    occ = np.zeros((9, 7, 5), dtype=np.int32)
    occ[0, 0, 0] = 1
    om = OracleMap(origin, res, size)
    gm = _gpu_map(origin, res, size)
    om.set_occupancy_buffer(occ), gm.upload_occupancy(occ)
    om.update_esdf(), gm.update_esdf()
    d = gm.distance()
    assert np.array_equal(d, om.distance())
    assert d.max() == 10000.0 and (d == 10000.0).sum() > 10 and d.reshape(9, 7, 5)[0, 0, 0] == 0.0


def test_edt_update_twice_and_reset(gtop_lib):
    """updateESDF3d takes min(new, old) (:358-360): a second obstacle set can only lower distances
    until resetBuffer is called."""
    origin, res, size = [0, 0, 0], 0.2, [6, 6, 2]
    a = np.array([[1.0, 1.0, 1.0]])
    b = np.array([[5.0, 5.0, 1.0]])
    om = OracleMap(origin, res, size)
    gm = _gpu_map(origin, res, size)
    for pts in (a, b):
        om.set_occupancy(pts), om.update_esdf()
        gm.set_occupancy(pts), gm.update_esdf()
        assert np.array_equal(gm.distance(), om.distance())
    om.reset(), gm.reset()
    om.set_occupancy(b), om.update_esdf()
    gm.set_occupancy(b), gm.update_esdf()
    assert np.array_equal(gm.distance(), om.distance())


def test_set_occupancy_drops_points_near_faces(gtop_lib):
    """isInMap margin of 1e-4 (:55-69): the z = 0 layer of the text_input recipe is rejected (H13)."""
    cfg = fx.load_cfg("text_input")
    gm = _gpu_map(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
    gm.reset()
    gm.set_occupancy(cfg["obstacles"])
    assert int(gm.occupancy().sum()) == 7650
    edge = np.array([[-10.0 + 5e-5, 0, 1], [0, 10.0 - 5e-5, 1], [0, 0, 5.0], [-10.0 + 2e-4, 0, 1]])
    gm.reset()
    gm.set_occupancy(edge)
    assert int(gm.occupancy().sum()) == 1


def test_trilinear_query_bit_exact(gtop_lib):
    cfg = fx.load_cfg("text_input")
    om = oracle_map_from_cfg(cfg)
    gm = _gpu_map(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
    gm.upload_distance(om.distance())
    rng = fx.Rng(99)
    n = 20000
    pos = cfg["map_origin"] + (rng.uniform((n, 3)) * 1.1 - 0.05) * cfg["map_size"]  # ~14 % out of map
    pos = pos.astype(np.float32).astype(np.float64)  # float-valued like the optimizer's samples (H1)
    # cell-boundary and face cases
    extra = np.array([[0.25, 0.25, 2.05], [0.05, -0.05, 0.05], [-9.9999, 0.0, 1.0], [9.9999, 9.9999, 4.9999],
                      [-9.95, -9.95, 0.05], [9.96, 9.96, 4.96], [-10.0, 0, 1], [0, 0, 5.0]])
    pos = np.concatenate([pos, extra])
    d_ref, g_ref, oob_ref = om.query(pos)
    d_gpu, g_gpu, oob_gpu = gm.query(pos)
    assert oob_gpu == oob_ref and oob_ref > 100
    assert np.array_equal(d_gpu, d_ref)
    assert np.array_equal(g_gpu, g_ref)


def test_fast_integer_sqrt_is_correctly_rounded_everywhere(gtop_lib):
    """The x pass replaces the library sqrt by a MUFU-seeded Goldschmidt step with an exact-residual correction for
    squared distances up to 3*8192^2; bit parity of the distance field rests on it being the IEEE value for EVERY
    integer in that range — checked exhaustively on the device."""
    import ctypes as C
    bad = C.c_int64(-1)
    assert gtop_lib.gtop_debug_check_sqrt(3 * 8192 * 8192, C.byref(bad)) == 0
    assert bad.value == 0


def _window_sequence(om, gm, steps):
    """apply the same call sequence to the oracle map and the GPU map, comparing both buffers after every step"""
    for k, (what, *args) in enumerate(steps):
        getattr(om, what)(*args), getattr(gm, what)(*args)
        assert np.array_equal(gm.occupancy(), om.occupancy()), "occupancy differs after step %d (%s)" % (k, what)
        d_gpu, d_ref = gm.distance(), om.distance()
        assert np.array_equal(d_gpu, d_ref), "distance differs after step %d (%s): %g" % (k, what, np.abs(d_gpu - d_ref).max())
        lo, hi = gm.update_range()
        olo, ohi = om.update_range()
        assert np.array_equal(lo, olo) and np.array_equal(hi, ohi)


@pytest.mark.parametrize("dims,res", [((40, 36, 20), 0.1), ((23, 17, 9), 0.25), ((150, 130, 140), 0.05)])
def test_windowed_update_and_reset_bit_exact(gtop_lib, dims, res):
    """setUpdateRange / resetBuffer(min,max) / updateESDF3d on a box (sdf_map.cpp:28-53, :244-264, :310-368): the GPU
    map follows the oracle bit for bit through full and windowed updates, and the window resets after each update."""
    origin = np.array([-2.0, 1.0, 0.0])
    size = np.array(dims, dtype=np.float64) * res
    rng = fx.Rng(11 + dims[0])
    pts = origin + 1e-3 + rng.uniform((60, 3)) * (size - 2e-3)
    more = origin + 1e-3 + rng.uniform((40, 3)) * (size - 2e-3)
    a = origin + size * np.array([0.2, 0.3, 0.1])
    b = origin + size * np.array([0.7, 0.9, 0.8])
    c = origin + size * np.array([0.5, -1.0, 0.45])   # clamped on y-
    d = origin + size * np.array([2.0, 0.6, 0.95])    # clamped on x+
    om, gm = OracleMap(origin, res, size), _gpu_map(origin, res, size)
    _window_sequence(om, gm, [
        ("set_occupancy", pts), ("update_esdf",),                        # whole grid
        ("set_occupancy", more), ("set_update_range", a, b), ("update_esdf",),   # only the box sees the new points
        ("reset_range", c, d), ("set_update_range", c, d), ("update_esdf",),     # local rebuild
        ("set_update_range", a, d), ("update_esdf",),
        ("update_esdf",),                                                # whole grid again (:366-367)
        ("reset",), ("set_occupancy", more), ("set_update_range", c, b), ("update_esdf",),  # window first after a reset
    ])


def test_windowed_update_empty_window_and_errors(gtop_lib):
    from grad_traj_optimization_b200._capi import GtopError
    origin, res, size = [0.0, 0.0, 0.0], 0.5, [8.0, 6.0, 4.0]
    om, gm = OracleMap(origin, res, size), _gpu_map(origin, res, size)
    _window_sequence(om, gm, [("set_occupancy", np.array([[1.1, 1.1, 0.6]])),
                              ("set_update_range", [4.0, 3.0, 2.0], [7.0, 5.0, 3.9]), ("update_esdf",)])  # no obstacle inside
    assert np.all(gm.distance() == 10000.0)
    for bad in (([9.0, 0, 0], [10.0, 1, 1]), ([3.0, 3.0, 3.0], [2.0, 4.0, 3.5])):  # outside the map / inverted
        with pytest.raises(GtopError):
            gm.set_update_range(*bad)
        with pytest.raises(GtopError):
            gm.reset_range(*bad)
    lo, hi = gm.update_range()
    assert np.array_equal(lo, [0, 0, 0]) and np.array_equal(hi, [15, 11, 7])  # a rejected call leaves the window alone


def test_windowed_update_keeps_the_gather_copy_in_step(gtop_lib):
    """The evaluation kernels gather from the xy-quad copy of the distance buffer, which a windowed update / reset only
    rebuilds around the box: evaluations on such a map equal those on a map that received the same distances whole."""
    from grad_traj_optimization_b200 import api
    cfg = fx.load_cfg("opti_node")
    origin, res, size = cfg["map_origin"], cfg["resolution"], cfg["map_size"]
    gm = _gpu_map(origin, res, size)
    gm.update(cfg["obstacles"])
    wp = np.asarray(cfg["waypoints"], dtype=np.float64)
    lo = wp.min(axis=0) - 0.7
    hi = wp.max(axis=0) + 0.4
    extra = wp[1:-1] + np.array([0.35, -0.2, 0.1])
    gm.reset_range(lo, wp.mean(axis=0))
    gm.set_occupancy(extra)
    gm.set_update_range(lo, hi)
    gm.update_esdf()
    ref = _gpu_map(origin, res, size)
    ref.upload_distance(gm.distance())
    ps = api.ProblemSet.from_paths(wp[None], api.make_params(cfg["params"]))
    x0, (lb, ub) = ps.x0(), ps.bounds()
    X = fx.multi_starts(x0[0], lb[0], ub[0], 64, seed=3)
    c1, g1, _ = ps.eval_host(gm, X, step=2, starts_per_problem=64)
    c2, g2, _ = ps.eval_host(ref, X, step=2, starts_per_problem=64)
    assert np.array_equal(c1, c2) and np.array_equal(g1, g2)
    c1, g1, _ = ps.eval_host(gm, X[:5], step=2, starts_per_problem=1)   # general kernel
    c2, g2, _ = ps.eval_host(ref, X[:5], step=2, starts_per_problem=1)
    assert np.array_equal(c1, c2) and np.array_equal(g1, g2)


@pytest.mark.parametrize("time", [-1.0, 0.0, 0.8, 2.5])
def test_dynamic_box_query_bit_exact(gtop_lib, time):
    """EDTEnvironment::evaluateEDTWithGrad / evaluateCoarseEDT (edt_environment.cpp:75-137) against the oracle
    restatement, including cells whose corner voxels lie outside the map (-1, not clamped)."""
    cfg = fx.load_cfg("opti_node")
    om = oracle_map_from_cfg(cfg)
    gm = _gpu_map(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
    gm.upload_distance(om.distance())
    origin, size = np.asarray(cfg["map_origin"], float), np.asarray(cfg["map_size"], float)
    rng = fx.Rng(31)
    pos = origin + 1e-3 + rng.uniform((4000, 3)) * (size - 2e-3)   # includes points in the border cells
    boxes = np.concatenate([origin + rng.uniform((12, 3)) * size, rng.uniform((12, 3)) - 0.5, 0.2 + rng.uniform((12, 3))], axis=1)
    d_ref, g_ref, c_ref = om.query_dynamic(pos, boxes, time)
    d, g, c = gm.query_dynamic(pos, boxes, time)
    assert np.array_equal(d, d_ref) and np.array_equal(g, g_ref) and np.array_equal(c, c_ref)
    d0, g0, _ = gm.query_dynamic(pos, np.zeros((0, 9)), time)   # no boxes: min with 1e7 = static field
    ds, _, _ = om.query_dynamic(pos, np.zeros((0, 9)), time)
    assert np.array_equal(d0, ds)


@pytest.mark.parametrize("name", ["text_input", "opti_node"])
def test_windowed_update_golden(gtop_lib, name):
    """committed golden digest of a windowed rebuild on the upstream fixtures (tests/golden/make_golden.py)"""
    import hashlib
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import window_case
    cfg = fx.load_cfg(name)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "fixtures_edt.npz"))
    gm = _gpu_map(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
    gm.update(cfg["obstacles"])
    d0 = gm.distance()
    lo, hi, extra = window_case(cfg)
    gm.reset_range(lo, hi), gm.set_occupancy(extra), gm.set_update_range(lo, hi), gm.update_esdf()
    d = gm.distance()
    assert np.array_equal(np.frombuffer(hashlib.sha256(d.tobytes()).digest(), dtype=np.uint8), gold[name + "_window_sha256"])
    assert int((d != d0).sum()) == int(gold[name + "_window_changed"])
