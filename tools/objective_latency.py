#!/usr/bin/env python
"""Latency of the B = 1 NLopt callback gtop_objective (the drop-in boundary) on the opti_node fixture."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from grad_traj_optimization_b200 import api, fixtures as fx
cfg = fx.load_cfg("opti_node")
gm = api.SdfMap(cfg["map_origin"], cfg["resolution"], cfg["map_size"]); gm.update(cfg["obstacles"])
ps = api.ProblemSet.from_paths(cfg["waypoints"], api.make_params(cfg["params"]))
obj = api.Objective(gm, ps, 0, step=2)
x = ps.x0()[0].copy(); g = np.empty_like(x)
for _ in range(200): obj(x, g)
t0 = time.perf_counter(); N = 3000
for _ in range(N): obj(x, g)
dt = (time.perf_counter() - t0) / N
print("gtop_objective (m=%d, n=%d): %.1f us per call (python ctypes overhead included)" % (ps.m, ps.n, dt * 1e6))
obj.close(); del gm, ps
