#!/usr/bin/env python
"""Measures the BASELINE.json configurations that bench.py does not headline (GPU box):
   edtA   : Map A EDT rebuild, warm, per-pass times and GB/s (25 B/voxel algorithmic)
   cfg2   : 1,024 distinct 7-waypoint problems (m=6), one start each
   cfg5   : m=30, 200 samples/segment, B=16,384
   cfg4   : Map B 2000x2000x400: EDT rebuild (box pillars), exactness spot-check vs the analytic box distance,
            4,096 distinct problems (m=10) gathering from the 25.6 GB (non L2-resident) grid
   python tools/config_bench.py [edtA cfg2 cfg5 cfg4]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PEAK = 6552.3


def time_eval(api, torch, gm, ps, X, starts, iters=30):
    B, n = X.shape
    stream = torch.cuda.Stream()
    Xd = [torch.from_numpy(X).cuda() for _ in range(4)]
    Gd = [torch.empty_like(Xd[0]) for _ in range(4)]
    Cd = [torch.empty(B, dtype=torch.float64, device="cuda") for _ in range(4)]
    for i in range(4):
        ps.eval_device(gm, Xd[i], Cd[i], Gd[i], starts_per_problem=starts, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(iters):
        ps.eval_device(gm, Xd[i % 4], Cd[i % 4], Gd[i % 4], starts_per_problem=starts, stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, Cd[0].cpu().numpy(), Gd[0].cpu().numpy()


def report(name, ms, B, m, ns, n, extra=None):
    bytes_eval = m * ns * 64 + 16 * n + 8
    gbs = B * bytes_eval / ms / 1e6
    out = {"config": name, "ms_per_batch": ms, "evals_per_s": B / ms * 1e3, "batch": B, "segments": m, "samples_per_segment": ns,
           "algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / PEAK}
    out.update(extra or {})
    print(json.dumps(out))
    return out


def main():
    which = sys.argv[1:] or ["edtA", "cfg2", "cfg5", "cfg4"]
    import torch
    from grad_traj_optimization_b200 import api, fixtures as fx
    gm = api.SdfMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
    obs = fx.map_a_obstacles()
    gm.update(obs)
    dist = gm.distance()
    if "edtA" in which:
        best = None
        for _ in range(5):
            gm.reset(), gm.set_occupancy(obs), gm.update_esdf()
            ms = gm.last_edt_ms()
            best = ms if best is None or ms[3] < best[3] else best
        nv = gm.nvox
        print(json.dumps({"config": "edtA 200x200x50", "ms_z_y_x_total": best, "algorithmic_GBps": nv * 25 / best[3] / 1e6,
                          "frac_of_measured_hbm": nv * 25 / best[3] / 1e6 / PEAK}))
    if "cfg2" in which:
        m, B = 6, 1024
        wps = np.stack([fx.random_waypoints(p, m, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"]) for p in range(B)])
        ps = api.ProblemSet.from_paths(wps, api.make_params(fx.PARAMS_OPTI))
        ms, _, _ = time_eval(api, torch, gm, ps, ps.x0(), 1)
        report("cfg2: 1,024 distinct 7-waypoint problems", ms, B, m, 30, ps.n)
    if "cfg5" in which:
        m, ns, sets, starts = 30, 200, 16, 1024
        wps = np.stack([fx.random_waypoints(p, m, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"]) for p in range(sets)])
        ps = api.ProblemSet.from_paths(wps, api.make_params(dict(fx.PARAMS_OPTI, samples_per_segment=ns)))
        x0, (lb, ub) = ps.x0(), ps.bounds()
        X = np.concatenate([fx.multi_starts(x0[p], lb[p], ub[p], starts, seed=7 + p) for p in range(sets)])
        ms, _, _ = time_eval(api, torch, gm, ps, X, starts, iters=10)
        report("cfg5: m=30, 200 samples/segment, B=16,384", ms, X.shape[0], m, ns, ps.n)
    del gm
    if "cfg4" in which:
        t0 = time.time()
        gb = api.SdfMap(fx.MAP_B["map_origin"], fx.MAP_B["resolution"], fx.MAP_B["map_size"])
        boxes = fx.box_pillars(3000, 4, gb.grid)
        best = None
        for _ in range(3):
            gb.reset(), gb.set_occupancy_boxes(boxes), gb.update_esdf()
            ms = gb.last_edt_ms()
            best = ms if best is None or ms[3] < best[3] else best
        nv = gb.nvox
        print(json.dumps({"config": "cfg4 EDT 2000x2000x400", "ms_z_y_x_total": best, "algorithmic_GBps": nv * 25 / best[3] / 1e6,
                          "frac_of_measured_hbm": nv * 25 / best[3] / 1e6 / PEAK, "setup_s": time.time() - t0}))
        # exactness spot-check: trilinear query at voxel centres returns the stored value
        rng = fx.Rng(11)
        vox = np.stack([np.floor(rng.uniform((20000,)) * g).astype(np.int64) for g in gb.grid], axis=1)
        vox = np.clip(vox, 1, np.array(gb.grid) - 2)
        centres = fx.MAP_B["map_origin"] + (vox + 0.5) * fx.MAP_B["resolution"]
        d_gpu, _, _ = gb.query(centres)
        d_ref = fx.boxes_exact_distance(boxes, vox, fx.MAP_B["resolution"])
        print(json.dumps({"config": "cfg4 EDT exactness", "voxels_checked": int(vox.shape[0]),
                          "max_abs_diff_vs_analytic_box_distance": float(np.abs(d_gpu - d_ref).max())}))
        m, B = 10, 4096
        free = fx.boxes_clearance_fn(boxes, fx.MAP_B["map_origin"], fx.MAP_B["resolution"], 0.5)
        wps = np.stack([fx.random_waypoints_free(p, m, free, 46.0, -45.0 + 80.0 * ((p * 0.6180339887) % 1.0), start_y_span=40.0)
                        for p in range(B)])
        ps = api.ProblemSet.from_paths(wps, api.make_params(fx.PARAMS_OPTI))
        ms, c, g = time_eval(api, torch, gb, ps, ps.x0(), 1)
        report("cfg4: 4,096 distinct problems on the 2000x2000x400 map", ms, B, m, 30, ps.n,
               {"finite": bool(np.isfinite(c).all() and np.isfinite(g).all()), "out_of_map": ps.take_out_of_map()})


if __name__ == "__main__":
    main()
