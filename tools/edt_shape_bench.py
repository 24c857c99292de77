#!/usr/bin/env python
"""EDT pass times for different grid shapes with the same row lengths (is the x pass bound by its 3.2 MB stride?).
   python tools/edt_shape_bench.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from grad_traj_optimization_b200 import api, fixtures as fx

for dims in [(2000, 2000, 400), (2000, 500, 400), (500, 2000, 400), (2000, 2000, 100), (1000, 1000, 400)]:
    res = 0.05
    gm = api.SdfMap([0, 0, 0], res, [d * res for d in dims])
    assert tuple(gm.grid) == dims, gm.grid
    boxes = fx.box_pillars(max(8, 3000 * dims[0] * dims[1] // 4000000), 4, gm.grid)
    best = None
    for _ in range(3):
        gm.reset(), gm.set_occupancy_boxes(boxes), gm.update_esdf()
        ms = gm.last_edt_ms()
        best = ms if best is None or ms[3] < best[3] else best
    nv = gm.nvox
    print(json.dumps({"grid": dims, "ms_z_y_x_total": [round(v, 3) for v in best],
                      "ns_per_kvoxel_z_y_x": [round(v * 1e6 / nv * 1e3, 2) for v in best[:3]],
                      "GBps": round(nv * 25 / best[3] / 1e6, 1)}))
    del gm
