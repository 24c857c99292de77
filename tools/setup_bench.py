#!/usr/bin/env python
"""Set-up cost of problem sets with one waypoint set per trajectory (configs 2 and 4): wall clock of
ProblemSet.from_paths (host set-up on all threads + record upload) next to one evaluation of the same batch.
   python tools/setup_bench.py [--m 10] [--sets 4096]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=10)
    ap.add_argument("--sets", type=int, default=4096)
    ap.add_argument("--device-setup", type=int, default=0)
    args = ap.parse_args()
    import torch
    from grad_traj_optimization_b200 import api, fixtures as fx
    gm = api.SdfMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
    gm.update(fx.map_a_obstacles())
    dist = gm.distance()
    base = [fx.random_waypoints(p, args.m, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"]) for p in range(16)]
    rng = np.random.default_rng(5)
    wps = np.stack([base[p % 16] + 0.05 * rng.standard_normal((args.m + 1, 3)) for p in range(args.sets)])
    prm = api.make_params(dict(fx.PARAMS_OPTI))
    out = {}
    for rep in range(3):
        t0 = time.perf_counter()
        ps = api.ProblemSet.from_paths(wps, prm, setup="device" if args.device_setup else "host")
        out["from_paths_ms"] = (time.perf_counter() - t0) * 1e3
    X = ps.x0()
    Xd = torch.from_numpy(X).cuda()
    Gd, Cd = torch.empty_like(Xd), torch.empty(X.shape[0], dtype=torch.float64, device="cuda")
    st = torch.cuda.Stream()
    for _ in range(3):
        ps.eval_device(gm, Xd, Cd, Gd, starts_per_problem=1, stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(20):
        ps.eval_device(gm, Xd, Cd, Gd, starts_per_problem=1, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    out.update(setup="device" if args.device_setup else "host", m=args.m, sets=args.sets, eval_ms=e0.elapsed_time(e1) / 20, host_threads=os.cpu_count())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
