"""Regenerates tests/golden/reference_run.npz from the REFERENCE'S OWN SOURCES — run from the repo root, in the
container that has /root/reference:

    make -C oracle reference && python tests/golden/make_reference_golden.py

oracle/_ref/libgtop_reference.so is src/sdf_map.cpp, src/qp_generator.cpp and src/grad_traj_optimizer.cpp of the
reference, compiled unmodified against the stand-in headers of oracle/shim/ (see oracle/ref_capi.cpp).  The vectors
stored here are outputs of that library on the reference's two fixtures; they travel to the GPU box (which has no
/root/reference) and pin both the oracle (tests/test_reference_build.py) and the CUDA path (tests/test_gpu_*.py):
  * the distance field after updateSDFMap (sha256 of the buffer, a sub-sampled slice, the maximum),
  * getDistWithGradTrilinear on a fixed point set (interior, faces, outside),
  * setPath products (segment times, L, R, Df, the initial derivatives),
  * costFunc (cost, gradient) on 16 multi-starts for OPT_FIRST_STEP and OPT_SECOND_STEP,
  * optimizeTrajectory(step 2) stopped after 40 evaluations (NLopt 2.5.0 LD_MMA, the vendored binary)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from grad_traj_optimization_b200 import fixtures as fx  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_run.npz")


def query_points(cfg, n=512, seed=11):
    """interior points, points on / next to every face, points outside the map"""
    rng = np.random.default_rng(seed)
    o, s = np.asarray(cfg["map_origin"], float), np.asarray(cfg["map_size"], float)
    p = o + s * rng.random((n, 3))
    k = n // 8
    for a in range(3):
        p[a * k:(a + 1) * k, a] = o[a] + rng.choice([1e-4, 1.0001e-4, 0.5 * cfg["resolution"], cfg["resolution"]], k)
        p[(3 + a) * k // 1:(3 + a) * k + k, a] = o[a] + s[a] - rng.choice([1e-4, 1.0001e-4, 0.4 * cfg["resolution"]], k)
    p[6 * k:7 * k] += s * rng.choice([-1.0, 1.0], (k, 3)) * (rng.random((k, 3)) < 0.4)
    return p


def fixture_starts(x0, lb, ub, nstart=16, seed=5):
    return fx.multi_starts(x0, lb, ub, nstart, seed=seed)


def main():
    out = {}
    for name in ("text_input", "opti_node"):
        cfg = fx.load_cfg(name)
        ro = rb.ReferenceOptimizer(cfg["params"])
        ro.init_map(cfg["map_size"], cfg["map_origin"], cfg["resolution"])
        ro.update_map(cfg["obstacles"])
        d = ro.distance()
        out[name + "_grid"] = np.array(d.shape)
        out[name + "_dist_sha256"] = np.frombuffer(hashlib.sha256(d.tobytes()).digest(), dtype=np.uint8)
        out[name + "_dist_max"] = np.array(d.max())
        out[name + "_dist_sub"] = d[::7, ::5, ::3].copy()
        rm = rb.ReferenceMap(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
        rm.reset(), rm.set_occupancy(cfg["obstacles"]), rm.update_esdf()
        q = query_points(cfg)
        qd, qg = rm.query(q)
        out[name + "_query_pos"], out[name + "_query_dist"], out[name + "_query_grad"] = q, qd, qg
        ro.set_path(cfg["waypoints"])
        out[name + "_seg_time"], out[name + "_L"], out[name + "_R"], out[name + "_Df"] = ro.seg_time(), ro.L(), ro.R(), ro.Df()
        x0 = ro.x0()
        out[name + "_x0"] = x0
        prm = cfg["params"]
        wp = np.asarray(cfg["waypoints"], float)
        lb, ub = x0.copy(), x0.copy()  # bounds as optimizeTrajectory sets them (:151-179)
        P = ro.num_dp
        for a in range(3):
            for i in range(P):
                c, h = (wp[i // 3 + 1, a], prm["bos"]) if i % 3 == 0 else (0.0, prm["vos"] if i % 3 == 1 else prm["aos"])
                lb[a * P + i], ub[a * P + i] = c - h, c + h
        X = fixture_starts(x0, lb, ub)
        out[name + "_X"] = X
        for step in (1, 2):
            c, g = ro.eval_batch(X, step)
            out["%s_cost%d" % (name, step)], out["%s_grad%d" % (name, step)] = c, g
        xs, ev = ro.optimize(2, 40)
        out[name + "_solve40_x"], out[name + "_solve40_evals"] = xs, np.array(ev)
        out[name + "_solve40_coeff"] = ro.coeff()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
