#!/usr/bin/env python
"""Which direction limits the in-place (zero-copy) host path?  k_eval_mma on the config-3 batch with x and / or the
outputs in pinned host memory, the rest in device memory (GPU box).   python tools/zero_copy_probe.py"""
import os, sys, json
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
Gh, Ch = torch.empty((B, n), dtype=torch.float64).pin_memory(), torch.empty((B,), dtype=torch.float64).pin_memory()
Xd, Gd, Cd = Xh.cuda(), torch.empty((B, n), dtype=torch.float64, device="cuda"), torch.empty((B,), dtype=torch.float64, device="cuda")
stream = torch.cuda.Stream()
out = {}
for name, (x, c, g) in {"all_device": (Xd, Cd, Gd), "x_from_host": (Xh, Cd, Gd), "outputs_to_host": (Xd, Ch, Gh),
                        "both": (Xh, Ch, Gh)}.items():
    for _ in range(3):
        ps.eval_device(gm, x, c, g, starts_per_problem=1024, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        ps.eval_device(gm, x, c, g, starts_per_problem=1024, stream=stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out[name] = {"ms": ms, "GBps_one_way": B * n * 8 / ms / 1e6}
print(json.dumps(out))
