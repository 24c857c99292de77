#!/usr/bin/env python
"""EDT rebuild timing (GPU box).  python tools/edt_bench.py [A|B] [reps]
   A: Map A 200x200x50 (random pillars through setOccupancy); B: Map B 2000x2000x400 (3,000 box pillars)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from grad_traj_optimization_b200 import api, fixtures as fx
which = sys.argv[1] if len(sys.argv) > 1 else "B"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if which == "A":
    gm = api.SdfMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
    obs = fx.map_a_obstacles()
    fill = lambda: gm.set_occupancy(obs)
else:
    gm = api.SdfMap(fx.MAP_B["map_origin"], fx.MAP_B["resolution"], fx.MAP_B["map_size"])
    boxes = fx.box_pillars(3000, 4, gm.grid)
    fill = lambda: gm.set_occupancy_boxes(boxes)
best = None
for _ in range(reps):
    gm.reset(), fill(), gm.update_esdf()
    ms = gm.last_edt_ms()
    best = ms if best is None or ms[3] < best[3] else best
nv = gm.nvox
print(json.dumps({"map": which, "grid": list(gm.grid), "ms_z_y_x_total": [round(v, 4) for v in best],
                  "algorithmic_GBps": nv * 25 / best[3] / 1e6, "frac_of_measured_hbm": nv * 25 / best[3] / 1e6 / 6552.3}))
del gm

# windowed rebuild (setUpdateRange + updateESDF3d on a box): a 10 m x 10 m x 5 m box of Map B, wall clock of the whole
# call (gather + three passes + merge + box-local quad rebuild + synchronize)
if which == "B" and os.environ.get("EDT_BENCH_WINDOW", "1") == "1":
    import time
    gm = api.SdfMap(fx.MAP_B["map_origin"], fx.MAP_B["resolution"], fx.MAP_B["map_size"])
    gm.set_occupancy_boxes(boxes), gm.update_esdf()
    full_ms = gm.last_edt_ms()[3]
    lo, hi = np.array([-5.0, -5.0, 0.0]), np.array([5.0, 5.0, 5.0])
    out = []
    for _ in range(reps + 1):
        gm.reset_range(lo, hi), gm.set_occupancy_boxes(boxes), gm.set_update_range(lo, hi)
        wlo, whi = gm.update_range()
        t0 = time.perf_counter()
        gm.update_esdf()
        out.append((time.perf_counter() - t0) * 1e3)
    wv = int(np.prod(whi - wlo + 1))
    print(json.dumps({"map": "B window", "window_voxels": wv, "window_grid": (whi - wlo + 1).tolist(),
                      "wall_ms_best": round(min(out[1:]), 4), "kernels_ms": round(gm.last_edt_ms()[3], 4),
                      "full_rebuild_kernels_ms": round(full_ms, 4)}))
