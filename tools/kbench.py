#!/usr/bin/env python
"""Quick kernel iteration loop (GPU box): times the evaluation kernel on the config-3 batch and checks a slice against the
oracle.  Not the judged benchmark (that is bench.py).   python tools/kbench.py [--m 10] [--ns 30] [--sets 64] [--starts 1024]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=10)
    ap.add_argument("--ns", type=int, default=30)
    ap.add_argument("--sets", type=int, default=64)
    ap.add_argument("--starts", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--check", type=int, default=256)
    ap.add_argument("--step", type=int, default=2, help="1 = OPT_FIRST_STEP (collision only), 2 = OPT_SECOND_STEP")
    ap.add_argument("--dump", default="", help="save cost / gradient of the first buffer set to this .npz")
    ap.add_argument("--bufs", type=int, default=4, help="rotating X / grad / cost buffer sets (working set > L2)")
    args = ap.parse_args()
    import torch
    from grad_traj_optimization_b200 import api, fixtures as fx
    gm = api.SdfMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
    gm.update(fx.map_a_obstacles())
    dist = gm.distance()
    wps = np.stack([fx.random_waypoints(p, args.m, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"]) for p in range(args.sets)])
    prm = dict(fx.PARAMS_OPTI, samples_per_segment=args.ns)
    ps = api.ProblemSet.from_paths(wps, api.make_params(prm))
    x0, (lb, ub) = ps.x0(), ps.bounds()
    X = np.concatenate([fx.multi_starts(x0[p], lb[p], ub[p], args.starts, seed=7 + p) for p in range(args.sets)])
    B, n = X.shape
    Xd = [torch.from_numpy(X).cuda() for _ in range(args.bufs)]
    Gd = [torch.empty_like(Xd[0]) for _ in range(args.bufs)]
    Cd = [torch.empty(B, dtype=torch.float64, device="cuda") for _ in range(args.bufs)]
    stream = torch.cuda.Stream()  # explicit stream: a NULL stream means "the handle's own stream" to the library
    st = stream.cuda_stream
    for i in range(5):
        ps.eval_device(gm, Xd[i % args.bufs], Cd[i % args.bufs], Gd[i % args.bufs], starts_per_problem=args.starts, stream=st, step=args.step)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.iters):
        ps.eval_device(gm, Xd[i % args.bufs], Cd[i % args.bufs], Gd[i % args.bufs], starts_per_problem=args.starts, stream=st, step=args.step)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    bytes_eval = args.m * args.ns * 64 + 16 * n + 8
    print("m=%d ns=%d B=%d: %.4f ms/batch, %.3e evals/s, %.1f GB/s algorithmic = %.1f%% of 6552.3" % (
        args.m, args.ns, B, ms, B / ms * 1e3, B * bytes_eval / ms / 1e6, B * bytes_eval / ms / 1e6 / 65.523))
    if args.dump:
        np.savez(args.dump, cost=Cd[0].cpu().numpy(), grad=Gd[0].cpu().numpy())
    if args.check:
        from common import oracle_problem, rel_cost_err, rel_grad_err
        from oracle import OracleMap
        om = OracleMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
        om.set_distance(dist)
        c, g = Cd[0].cpu().numpy(), Gd[0].cpu().numpy()
        ec = eg = 0.0
        full = args.check >= args.starts  # --check >= starts: every trajectory of every waypoint set
        for p in (range(args.sets) if full else (0, args.sets - 1)):
            opt = oracle_problem(prm, om, wps[p])
            sl = slice(p * args.starts, p * args.starts + min(args.check, args.starts))
            cr, gr, _, _ = opt.eval_batch(X[sl], args.step)
            ec, eg = max(ec, rel_cost_err(c[sl], cr).max()), max(eg, rel_grad_err(g[sl], gr).max())
        print("parity vs oracle%s: cost %.2e grad %.2e %s" % (" (all %d trajectories)" % B if full else "", ec, eg,
                                                             "OK" if max(ec, eg) <= 1e-9 else "FAIL"))


if __name__ == "__main__":
    main()
