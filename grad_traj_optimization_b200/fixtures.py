"""Fixture and synthetic-workload generators (host side, numpy only).

* `load_cfg` reads configs/*.cfg — the reference's demo values re-entered as data
  (launch/text_input.launch, src/example_text_input.cpp:28-70, src/opti_node.cpp:61-99,
  launch/opti_node.launch) — and expands the obstacle recipes into points exactly as the demos do
  (float z accumulation for the pillars, double accumulation for opti_node's nested loops).
* `map_a_*`, `random_waypoints`, `multi_starts` build the BASELINE.json workloads
  (random.launch-style 20x20x5 m @0.1 m map after src/example_click.cpp:59-118, 7/11/31-waypoint
  paths, multi-start batches) from a fixed counter-based RNG, so every rank / test / bench run sees
  the same inputs without shipping data files.
"""
import os

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG_DIR = os.path.join(_REPO, "configs")


# ------------------------------------------------------------------------------------------------
# configs
# ------------------------------------------------------------------------------------------------
def _loop_vals(start, end, step):
    """for (double v = start; v <= end (step>0) | v >= end (step<0); v += step)"""
    out, v = [], float(start)
    if step > 0:
        while v <= end:
            out.append(v)
            v = v + step
    else:
        while v >= end:
            out.append(v)
            v = v + step
    return out


def pillar_points(centres_xy, resolution, half_width=2, z_max=3.5):
    """example_text_input.cpp:60-70 / example_click.cpp:108-118: for(float z=0; z<z_max; z+=res)
    for each obstacle, (2*th+1)^2 points at centre + (m,n)*res."""
    zs, z = [], np.float32(0.0)
    while z < z_max:
        zs.append(float(z))
        z = np.float32(np.float64(z) + resolution)  # float += double
    c = np.asarray(centres_xy, dtype=np.float64).reshape(-1, 2)
    offs = np.arange(-half_width, half_width + 1, dtype=np.float64) * resolution
    pts = np.empty((len(zs), c.shape[0], offs.size, offs.size, 3))
    pts[..., 0] = (c[:, 0][None, :, None, None] + offs[None, None, :, None])
    pts[..., 1] = (c[:, 1][None, :, None, None] + offs[None, None, None, :])
    pts[..., 2] = np.asarray(zs)[:, None, None, None]
    return pts.reshape(-1, 3)


def load_cfg(name):
    """Parse configs/<name>.cfg -> dict(map_size, map_origin, resolution, obstacles[N,3], waypoints[K,3], params)."""
    path = name if os.path.exists(name) else os.path.join(CONFIG_DIR, name + ".cfg")
    cfg = {"params": {}, "pillars": [], "loop_boxes": [], "waypoints": []}
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if not line:
                continue
            key, *vals = line.split()
            if key in ("map_size", "map_origin"):
                cfg[key] = np.array([float(v) for v in vals])
            elif key in ("resolution", "pillar_z_max"):
                cfg[key] = float(vals[0])
            elif key == "pillar_half_width":
                cfg[key] = int(vals[0])
            elif key == "pillar":
                cfg["pillars"].append([float(v) for v in vals])
            elif key == "loop_box":
                cfg["loop_boxes"].append([float(v) for v in vals])
            elif key == "waypoint":
                cfg["waypoints"].append([float(v) for v in vals])
            elif key == "param":
                cfg["params"][vals[0]] = float(vals[1])
            else:
                raise ValueError("%s: unknown key %r" % (path, key))
    pts = []
    if cfg["pillars"]:
        pts.append(pillar_points(cfg["pillars"], cfg["resolution"], cfg.get("pillar_half_width", 2),
                                 cfg.get("pillar_z_max", 3.5)))
    for b in cfg["loop_boxes"]:
        xs, ys, zs = _loop_vals(*b[0:3]), _loop_vals(*b[3:6]), _loop_vals(*b[6:9])
        g = np.array([[x, y, z] for x in xs for y in ys for z in zs], dtype=np.float64)
        pts.append(g.reshape(-1, 3))
    cfg["obstacles"] = np.concatenate(pts, axis=0) if pts else np.zeros((0, 3))
    cfg["waypoints"] = np.asarray(cfg["waypoints"], dtype=np.float64)
    return cfg


# ------------------------------------------------------------------------------------------------
# deterministic counter-based RNG (splitmix64), identical on every platform / numpy version
# ------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


class Rng:
    """Stream `seed`: u(i) = splitmix64(splitmix64(seed) + i); uniform in [0,1) from the top 53 bits."""

    def __init__(self, seed):
        self.base = _splitmix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF))
        self.ctr = 0

    def uniform(self, size=None):
        n = 1 if size is None else int(np.prod(size))
        with np.errstate(over="ignore"):
            idx = self.base + (np.arange(self.ctr, self.ctr + n, dtype=np.uint64))
        self.ctr += n
        u = (_splitmix64(idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        return float(u[0]) if size is None else u.reshape(size)

    def normal(self, size):
        n = int(np.prod(size))
        u = self.uniform((2, (n + 1)))
        r = np.sqrt(-2.0 * np.log(1.0 - u[0]))  # Box-Muller; 1-u in (0,1]
        z = r * np.cos(2.0 * np.pi * u[1])
        return z[:n].reshape(size)


# ------------------------------------------------------------------------------------------------
# Map A: random.launch-style map (example_click.cpp:59-118), 20x20x5 m @0.1 m
# ------------------------------------------------------------------------------------------------
MAP_A = dict(map_size=np.array([20.0, 20.0, 5.0]), map_origin=np.array([-10.0, -10.0, 0.0]), resolution=0.1)
# opti_node.launch parameters, used for the synthetic batches
PARAMS_OPTI = dict(ws=1.0, wc=5.0, alpha=10.0, d0=0.8, r=0.5, mean_v=1.8, init_time=0.3, bos=3.0, vos=8.0, aos=10.0,
                   alpha_v=0.0, v0=2.5, r_v=1.5, alpha_a=0.0, a0=3.5, r_a=1.5, alg=24, time_limit_1=0.4,
                   time_limit_2=0.1, dt=0.2)


def map_a_pillar_centres(n_pillars=120, seed=20170924, half_extent=8.0, min_sep=1.2):
    rng = Rng(seed)
    c = []
    fails = 0
    while len(c) < n_pillars and fails < 10000:
        p = -half_extent + 2.0 * half_extent * rng.uniform((2,))
        if all((p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2 > min_sep ** 2 for q in c):
            c.append(p)
            fails = 0
        else:
            fails += 1
    return np.asarray(c)


def map_a_obstacles(n_pillars=120, seed=20170924):
    return pillar_points(map_a_pillar_centres(n_pillars, seed), MAP_A["resolution"], 2, 3.5)


def trilinear_free(dist, origin, res, p, clearance):
    """nearest-voxel clearance test on a host copy of the distance grid (workload shaping only)."""
    idx = np.floor((np.asarray(p) - origin) / res).astype(int)
    if np.any(idx < 0) or np.any(idx >= np.array(dist.shape)):
        return False
    return dist[idx[0], idx[1], idx[2]] >= clearance


def random_waypoints(problem_id, m, dist, origin, res, xy_limit=8.5, clearance=0.5, step_len=(1.2, 2.5),
                     jitter_deg=40.0, z_mid=2.0, z_amp=0.5, start_x=-8.0):
    """m+1 waypoints: random walk of step U[1.2,2.5] m, heading jitter +-40 deg, z = 2 +- 0.5, each
    point >= `clearance` from obstacles and inside |x|,|y| <= xy_limit (turns around at the walls)."""
    rng = Rng(1000 + problem_id)
    pts = []
    for _ in range(200):  # first point
        p = np.array([start_x, -6.0 + 12.0 * rng.uniform(), z_mid + z_amp * (2.0 * rng.uniform() - 1.0)])
        if trilinear_free(dist, origin, res, p, clearance):
            break
    pts.append(p)
    heading = 0.0
    while len(pts) < m + 1:
        ok = False
        for attempt in range(400):
            spread = np.deg2rad(jitter_deg) if attempt < 40 else np.pi
            h = heading + spread * (2.0 * rng.uniform() - 1.0)
            ln = step_len[0] + (step_len[1] - step_len[0]) * rng.uniform()
            q = np.array([pts[-1][0] + ln * np.cos(h), pts[-1][1] + ln * np.sin(h),
                          z_mid + z_amp * (2.0 * rng.uniform() - 1.0)])
            if abs(q[0]) <= xy_limit and abs(q[1]) <= xy_limit and trilinear_free(dist, origin, res, q, clearance):
                ok = True
                break
        if not ok:
            raise RuntimeError("random_waypoints: could not place waypoint %d of problem %d" % (len(pts), problem_id))
        heading = h
        pts.append(q)
    return np.asarray(pts)


def multi_starts(x0, lb, ub, n_starts, seed=7, sigma=(0.3, 0.5, 0.5)):
    """x = clip(x0 + sigma o N(0,1), lb, ub); start 0 is x0 itself.  x layout [axis][wp][pos,vel,acc]."""
    n = x0.size
    sig = np.tile(np.asarray(sigma), n // 3)
    z = Rng(seed).normal((n_starts, n))
    X = np.clip(x0[None, :] + sig[None, :] * z, lb[None, :], ub[None, :])
    X[0] = x0
    return X


def straight_line_x0(waypoints, bos=3.0, vos=8.0, aos=10.0):
    """x0 / bounds of a waypoint set, numpy restatement of getInitialD + the NLopt box
    (qp_generator.cpp:407-451, grad_traj_optimizer.cpp:151-187) for workload generation."""
    w = np.asarray(waypoints, dtype=np.float64)
    m = w.shape[0] - 1
    P = 3 * (m - 1)
    x0 = np.zeros(3 * P)
    half = np.zeros(3 * P)
    for a in range(3):
        x0[a * P:(a + 1) * P:3] = w[1:m, a]
        half[a * P + 0:(a + 1) * P:3] = bos
        half[a * P + 1:(a + 1) * P:3] = vos
        half[a * P + 2:(a + 1) * P:3] = aos
    return x0, x0 - half, x0 + half


# ------------------------------------------------------------------------------------------------
# Map B (BASELINE configs[3]): 100x100x20 m @0.05 m = 2000x2000x400 voxels, 3,000 box pillars
# ------------------------------------------------------------------------------------------------
MAP_B = dict(map_size=np.array([100.0, 100.0, 20.0]), map_origin=np.array([-50.0, -50.0, 0.0]), resolution=0.05)


def box_pillars(n_pillars, seed, grid, half_width_vox=5, height_vox=(40, 300), margin_vox=100):
    """Voxel boxes [x0,y0,z0,x1,y1,z1] (inclusive) of `n_pillars` square pillars standing on the floor:
    (2*half+1)^2 footprint, centre uniform inside the grid minus a margin, height uniform in `height_vox`."""
    rng = Rng(seed)
    u = rng.uniform((n_pillars, 3))
    cx = (margin_vox + u[:, 0] * (grid[0] - 2 * margin_vox)).astype(np.int64)
    cy = (margin_vox + u[:, 1] * (grid[1] - 2 * margin_vox)).astype(np.int64)
    hz = (height_vox[0] + u[:, 2] * (height_vox[1] - height_vox[0])).astype(np.int64)
    b = np.stack([cx - half_width_vox, cy - half_width_vox, np.zeros_like(cx), cx + half_width_vox, cy + half_width_vox,
                  np.minimum(hz, grid[2] - 1)], axis=1)
    return b.astype(np.int32)


def boxes_exact_distance(boxes, vox, resolution):
    """Exact unsigned EDT value at voxel indices `vox` [N,3] for a map whose occupied set is the union of
    the voxel boxes: min over boxes of the integer squared distance to the box, then res*sqrt (bit-exact
    with updateESDF3d, sdf_map.cpp:356-360)."""
    vox = np.asarray(vox, dtype=np.int64)
    best = np.full(vox.shape[0], np.iinfo(np.int64).max)
    lo, hi = boxes[:, 0:3].astype(np.int64), boxes[:, 3:6].astype(np.int64)
    for i in range(boxes.shape[0]):
        d = np.maximum(np.maximum(lo[i] - vox, vox - hi[i]), 0)
        best = np.minimum(best, (d * d).sum(axis=1))
    return np.minimum(resolution * np.sqrt(best.astype(np.float64)), 10000.0)


def boxes_clearance_fn(boxes, origin, resolution, clearance):
    """is_free(p) for waypoint placement on a box map: horizontal distance to every pillar footprint
    >= clearance (pillars treated as infinitely tall)."""
    lo = origin[:2] + boxes[:, 0:2] * resolution
    hi = origin[:2] + (boxes[:, 3:5] + 1) * resolution

    def is_free(p):
        d = np.maximum(np.maximum(lo - p[:2], p[:2] - hi), 0.0)
        return bool(np.min((d * d).sum(axis=1)) >= clearance * clearance)

    return is_free


def random_waypoints_free(problem_id, m, is_free, xy_limit, start_x, start_y_span=6.0, step_len=(1.2, 2.5),
                          jitter_deg=40.0, z_mid=2.0, z_amp=0.5):
    """`random_waypoints` with a caller-supplied clearance predicate (used for the large box map)."""
    rng = Rng(1000 + problem_id)
    for _ in range(400):
        p = np.array([start_x, -start_y_span + 2 * start_y_span * rng.uniform(), z_mid + z_amp * (2.0 * rng.uniform() - 1.0)])
        if is_free(p):
            break
    pts, heading = [p], 0.0
    while len(pts) < m + 1:
        for attempt in range(400):
            spread = np.deg2rad(jitter_deg) if attempt < 40 else np.pi
            h = heading + spread * (2.0 * rng.uniform() - 1.0)
            ln = step_len[0] + (step_len[1] - step_len[0]) * rng.uniform()
            q = np.array([pts[-1][0] + ln * np.cos(h), pts[-1][1] + ln * np.sin(h), z_mid + z_amp * (2.0 * rng.uniform() - 1.0)])
            if abs(q[0]) <= xy_limit and abs(q[1]) <= xy_limit and is_free(q):
                break
        else:
            raise RuntimeError("random_waypoints_free: could not place waypoint %d of problem %d" % (len(pts), problem_id))
        heading = h
        pts.append(q)
    return np.asarray(pts)
