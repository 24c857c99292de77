"""GPU voxel-map kernels vs the CPU oracle: occupancy scatter, exact EDT (bit-exact), trilinear query.

Reference behaviour under test: SDFMap::setOccupancy / updateESDF3d / getDistWithGradTrilinear
(src/sdf_map.cpp:80-99, 310-368, 185-242 upstream)."""
import numpy as np
import pytest

from common import oracle_map_from_cfg
from grad_traj_optimization_b200 import fixtures as fx
from oracle import OracleMap

pytestmark = pytest.mark.gpu


def _gpu_map(origin, res, size):
    from grad_traj_optimization_b200.api import SdfMap
    return SdfMap(origin, res, size)


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


def test_edt_empty_map_is_10000_everywhere(gtop_lib):
    gm = _gpu_map([0, 0, 0], 0.5, [8, 6, 4])
    gm.reset()
    gm.update_esdf()
    assert np.all(gm.distance() == 10000.0)  # sdf_map.cpp:22,358-360


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
