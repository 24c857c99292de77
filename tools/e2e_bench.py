#!/usr/bin/env python
"""End-to-end (host buffers) timing of gtop_eval_batch_host on the config-3 batch; GTOP_HOST_CHUNKS selects the
pipeline depth.   python tools/e2e_bench.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from grad_traj_optimization_b200 import api, fixtures as fx
gm = api.SdfMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
gm.update(fx.map_a_obstacles())
dist = gm.distance()
wps = np.stack([fx.random_waypoints(p, 10, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"]) for p in range(64)])
ps = api.ProblemSet.from_paths(wps, api.make_params(fx.PARAMS_OPTI))
x0, (lb, ub) = ps.x0(), ps.bounds()
X = np.concatenate([fx.multi_starts(x0[p], lb[p], ub[p], 1024, seed=7 + p) for p in range(64)])
B, n = X.shape
Xh = torch.from_numpy(X).pin_memory()
Gh = torch.empty((B, n), dtype=torch.float64).pin_memory()
Ch = torch.empty((B,), dtype=torch.float64).pin_memory()
for _ in range(3):
    ps.eval_host(gm, Xh, step=2, starts_per_problem=1024, cost=Ch, grad=Gh)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30):
    ps.eval_host(gm, Xh, step=2, starts_per_problem=1024, cost=Ch, grad=Gh)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 30 * 1e3
print("chunks=%s: %.3f ms per 65,536 batch end to end, %.3e evals/s, PCIe %.1f GB/s each way" % (
    os.environ.get("GTOP_HOST_CHUNKS", "default"), ms, B / ms * 1e3, B * n * 8 / ms / 1e6))
