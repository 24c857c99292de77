#!/usr/bin/env python
"""Timing of the post-hoc metrics kernel on the config-3 batch (65,536 trajectories x 10 segments).
   python tools/metrics_bench.py"""
import os, sys
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
Xd = torch.from_numpy(X).cuda()
out = torch.empty(X.shape[0], 8, dtype=torch.float64, device="cuda")
st = torch.cuda.Stream()
for _ in range(3):
    ps.traj_metrics_device(Xd, out, starts_per_problem=1024, stream=st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(20):
    ps.traj_metrics_device(Xd, out, starts_per_problem=1024, stream=st.cuda_stream)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
m = out.cpu().numpy()
print("traj metrics: %.3f ms per %d trajectories (%.2e trajectories/s); median length %.3f m, time %.3f s, ~%d points each"
      % (ms, X.shape[0], X.shape[0] / ms * 1e3, np.median(m[:, 1]), np.median(m[:, 0]), int(np.median(m[:, 0]) / 0.01)))
