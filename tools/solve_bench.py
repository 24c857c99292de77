#!/usr/bin/env python
"""Throughput of the batched on-device optimiser loop (gtop_solve_batch_device) on the config-3 batch:
65,536 multi-starts x `maxeval` MMA evaluations, x resident on the device.   python tools/solve_bench.py [maxeval]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from grad_traj_optimization_b200 import api, fixtures as fx
maxeval = int(sys.argv[1]) if len(sys.argv) > 1 else 100
gm = api.SdfMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
gm.update(fx.map_a_obstacles())
dist = gm.distance()
wps = np.stack([fx.random_waypoints(p, 10, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"]) for p in range(64)])
ps = api.ProblemSet.from_paths(wps, api.make_params(fx.PARAMS_OPTI))
x0, (lb, ub) = ps.x0(), ps.bounds()
X = np.concatenate([fx.multi_starts(x0[p], lb[p], ub[p], 1024, seed=7 + p) for p in range(64)])
B, n = X.shape
stream = torch.cuda.Stream()
Xd = torch.from_numpy(X).cuda()
minf = torch.empty(B, dtype=torch.float64, device="cuda")
c0 = torch.empty(B, dtype=torch.float64, device="cuda")
g0 = torch.empty_like(Xd)
ps.eval_device(gm, Xd, c0, g0, starts_per_problem=1024, stream=stream.cuda_stream)
work = Xd.clone()
ps.solve_device(gm, work, minf, maxeval=5, starts_per_problem=1024, stream=stream.cuda_stream)  # warm-up
torch.cuda.synchronize()
work.copy_(Xd)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(stream)
ps.solve_device(gm, work, minf, maxeval=maxeval, starts_per_problem=1024, stream=stream.cuda_stream)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
f0, f1 = c0.cpu().numpy(), minf.cpu().numpy()
print(json.dumps({"solve": "65,536 multi-starts x %d MMA evaluations on device" % maxeval, "ms_total": ms,
                  "ms_per_evaluation_step": ms / maxeval, "solves_per_s": B / ms * 1e3, "evals_per_s": B * maxeval / ms * 1e3,
                  "median_cost_start": float(np.median(f0)), "median_cost_end": float(np.median(f1)),
                  "all_improved": bool(np.all(f1 <= f0 + 1e-9))}))
del gm, ps
