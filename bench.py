#!/usr/bin/env python
"""bench.py — headline measurement of the GTOP inner loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): Map A — 20x20x5 m
@0.1 m random-pillar map — 64 waypoint sets x 1,024 multi-starts = 65,536 trajectories x 10 segments
x 30 samples per GPU.  One "step" = one batched cost+gradient evaluation of the whole batch.
`value` = evaluations/s with inputs resident in HBM (CUDA events, max over ranks); `e2e` = the same
through the reference-facing C ABI call gtop_eval_batch_host with pinned HOST buffers (H2D of x and D2H
of cost+grad inside the timed region).  Multi-GPU: one process per GPU (torchrun), the batch is
partitioned, the grid replicated, no collective on the data path (only the timing all-reduce).

--impl reference times the CPU port of the reference objective (oracle/, the checker) on all host
cores of rank 0 on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_SEG, N_SETS, N_STARTS, N_SAMP = 10, 64, 1024, 30
METRIC = "traj_cost_grad_evals_per_s"
UNIT = "evals/s"


def algorithmic_bytes_per_eval(m, ns, n, voxel_bytes=8):
    """8 trilinear corners per sample + read x + write grad + write cost."""
    return m * ns * 8 * voxel_bytes + 8 * n + 8 * n + 8


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU while the timed regions run (pynvml, 25 ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.active = index, [], set(), False, False
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.025)

    def summary(self):
        if not self.samples:
            # fall back to one nvidia-smi sample
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout
                a, b = [float(v) for v in out.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 0}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz), "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic_bytes():
    """dram__bytes_read+write per launch of the dominant kernel from the committed ncu capture (or None)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d.get("k_eval_mma_dram_bytes_per_launch", d.get("k_eval_traj_dram_bytes_per_launch"))
        except Exception:
            return None
    return None


def build_workload(api, fx, device, rank, per_gpu_batch):
    """Map A on this GPU (GPU EDT), 64 waypoint sets for this rank, multi-start X for its shard."""
    gm = api.SdfMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"], device=device)
    gm.update(fx.map_a_obstacles())
    dist = gm.distance()
    n_sets = max(1, per_gpu_batch // N_STARTS)
    wps = np.stack([fx.random_waypoints(rank * N_SETS + p, M_SEG, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"])
                    for p in range(n_sets)])
    prm = api.make_params(fx.PARAMS_OPTI, samples_per_segment=N_SAMP)
    ps = api.ProblemSet.from_paths(wps, prm, device=device)
    x0, (lb, ub) = ps.x0(), ps.bounds()
    starts = per_gpu_batch // n_sets
    X = np.concatenate([fx.multi_starts(x0[p], lb[p], ub[p], starts, seed=7 + rank * N_SETS + p) for p in range(n_sets)])
    return gm, ps, X, starts, wps


def run_ours(args):
    import torch
    import torch.distributed as dist
    from grad_traj_optimization_b200 import api, fixtures as fx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if api.lib().gtop_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the GPU arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the ONE JSON line: communicator creation prints "NCCL version ..." on stdout, so stdout
        # points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    n_gpus = world

    full = N_SETS * N_STARTS
    per_gpu = full if args.scaling == "weak" else full // n_gpus
    gm, ps, X, starts, _ = build_workload(api, fx, local, rank if args.scaling == "weak" else 0, full)
    if args.scaling == "strong":  # contiguous block b*G/B of the one global batch
        lo = rank * per_gpu
        X, traj_offset = X[lo:lo + per_gpu], lo
    else:
        traj_offset = 0
    B, n = X.shape
    stream = torch.cuda.Stream()
    NBUF = 4  # rotating buffer sets: 4 x (X + grad) = 340 MB > 126 MB L2
    Xd = [torch.from_numpy(X).cuda() for _ in range(NBUF)]
    Gd = [torch.empty((B, n), dtype=torch.float64, device="cuda") for _ in range(NBUF)]
    Cd = [torch.empty((B,), dtype=torch.float64, device="cuda") for _ in range(NBUF)]
    torch.cuda.synchronize()

    def step_device(i):
        j = i % NBUF
        ps.eval_device(gm, Xd[j], Cd[j], Gd[j], step=2, traj_offset=traj_offset, starts_per_problem=starts,
                       stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-resident arm -------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step_device(i)
    barrier()
    launches0 = api.lib().gtop_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active = True
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(args.steps):
            step_device(i)
        e1.record(stream)
    barrier()
    sampler.active = False
    ms_total = e0.elapsed_time(e1)
    launches = api.lib().gtop_kernel_launch_count() - launches0
    oob = ps.take_out_of_map()

    # ---- end-to-end arm: C ABI with pinned host buffers ----------------------------------------
    Xh = torch.from_numpy(X).pin_memory()
    Gh = torch.empty((B, n), dtype=torch.float64).pin_memory()
    Ch = torch.empty((B,), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        ps.eval_host(gm, Xh, step=2, traj_offset=traj_offset, starts_per_problem=starts, cost=Ch, grad=Gh)
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ps.eval_host(gm, Xh, step=2, traj_offset=traj_offset, starts_per_problem=starts, cost=Ch, grad=Gh)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.active = False
    sampler.stop_flag = True
    # device result and host-path result of the same inputs must be identical
    same = bool(np.array_equal(Ch.numpy(), Cd[0].cpu().numpy()))

    # ---- optimiser loop on device: the same batch through 100 lock-step LD_MMA evaluations ------
    solve = None
    if not args.no_solve:
        try:
            work = Xd[0].clone()
            minf = torch.empty((B,), dtype=torch.float64, device="cuda")
            ps.solve_device(gm, work, minf, step=2, maxeval=5, traj_offset=traj_offset, starts_per_problem=starts,
                            stream=stream.cuda_stream)  # warm-up
            work.copy_(Xd[0])
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            ps.solve_device(gm, work, minf, step=2, maxeval=100, traj_offset=traj_offset, starts_per_problem=starts,
                            stream=stream.cuda_stream)
            s1.record(stream)
            barrier()
            solve_ms = s0.elapsed_time(s1)
            solve = {"maxeval": 100, "ms": solve_ms, "solves_per_s_per_gpu": B / solve_ms * 1e3,
                     "median_cost_start": float(torch.median(Cd[0]).item()), "median_cost_end": float(torch.median(minf).item())}
        except Exception as exc:  # the headline measurement above does not depend on this leg
            solve = {"error": str(exc)}

    # ---- FP64 ceilings measured in this run ------------------------------------------------
    fp64 = None
    if rank == 0:
        try:
            import ctypes as C
            dfma, dmma = C.c_double(0), C.c_double(0)
            if api.lib().gtop_measure_fp64_peak(local, C.byref(dfma), C.byref(dmma)) == 0:
                fp64 = {"dfma_peak_tflops": dfma.value, "dmma_peak_tflops": dmma.value}
        except Exception as exc:
            fp64 = {"error": str(exc)}

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        ms_step = ms_total / args.steps
        total_evals = B * n_gpus
        value = total_evals / (ms_step * 1e-3)
        bytes_eval = algorithmic_bytes_per_eval(M_SEG, N_SAMP, n)
        peak, peak_kind = measured_peak_gbs()
        achieved = B * bytes_eval / (ms_step * 1e-3) / 1e9  # per-GPU kernel: one launch = B evaluations
        cpu = cpu_baseline(args, X[:min(B, 4096)], starts) if not args.no_cpu_baseline else None
        if fp64 and "dfma_peak_tflops" in fp64:
            # executed FP64 work of k_eval_mma per trajectory-segment (ncu, profiles/r1_k_eval_mma_ncu.txt): 99.2
            # DADD/DMUL + 32.5 DFMA warp instructions (32 lanes) + 6.0 DMMA (512 flops each)
            flops_unit = (52.3 + 46.9) * 32 + 32.5 * 64 + 6.0 * 512
            tf = flops_unit * M_SEG * B / (ms_step * 1e-3) / 1e12
            fp64.update({"kernel_executed_tflops": tf, "frac_of_dfma_peak": tf / fp64["dfma_peak_tflops"],
                         "fp64_unit_busy_ncu": "vector 44.5 % + DMMA 16.2 % of cycles (profiles/r1_k_eval_mma_ncu.txt)"})
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[2]: Map A 20x20x5m@0.1m (200x200x50 fp64 SDF), 64 waypoint sets x 1024 "
                                   "multi-starts, m=10 segments x 30 samples, step=2 (opti_node.launch params)",
                       "per_gpu_batch": B, "global_batch": total_evals, "segments": M_SEG, "samples_per_segment": N_SAMP,
                       "parallelism": "batch partitioned across %d GPU(s), SDF replicated, no collective" % n_gpus,
                       "l2": "4 rotating X/grad buffer sets (340 MB > 126 MB L2); the 16 MB SDF grid is L2-resident by "
                             "construction", "out_of_map_samples_per_step": oob // max(1, args.steps + max(args.warmup, 3))},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_bytes(), "peak_kind": peak_kind, "kernel": "k_eval_mma",
                         "algorithmic_bytes_per_eval": bytes_eval, "kernel_ms": ms_step},
            "cpu_baseline": cpu,
            "e2e": {"value": total_evals / (e2e_ms * 1e-3 / e2e_steps), "unit": UNIT, "h2d_bytes_per_step": B * n * 8,
                    "d2h_bytes_per_step": B * n * 8 + B * 8, "steps": e2e_steps, "host_equals_device": same},
            "gpu_launches": int(launches),
            "solve_on_device": solve,
            "fp64": fp64,
            "clocks": sampler.summary(),
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _oracle_problems(X, starts):
    """Oracle contexts for the bounded CPU sample: Map A by the oracle's own EDT, the sample's waypoint sets."""
    from grad_traj_optimization_b200 import fixtures as fx
    from oracle import OracleMap, OracleOptimizer, OracleParams
    om = OracleMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
    om.reset()
    om.set_occupancy(fx.map_a_obstacles())
    om.update_esdf()
    dist = om.distance()
    n_sets = max(1, X.shape[0] // starts)
    prm = {"w_smooth": fx.PARAMS_OPTI["ws"], "w_collision": fx.PARAMS_OPTI["wc"], "samples_per_segment": N_SAMP,
           **{k: v for k, v in fx.PARAMS_OPTI.items() if k not in ("ws", "wc", "alg", "dt")}}
    opts = []
    for p in range(n_sets):
        w = fx.random_waypoints(p, M_SEG, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"])
        o = OracleOptimizer(OracleParams.from_dict(prm), om)
        o.set_path(w)
        opts.append(o)
    return opts


def cpu_baseline(args, Xs, starts, seconds=10.0):
    """CPU port of the reference objective (oracle/) on all host cores, on the first <=4,096 trajectories of
    the workload, repeated for ~10 s.  A reported baseline, not the optimisation target."""
    from oracle import binding
    opts = _oracle_problems(Xs, starts)
    threads = binding.hardware_threads()
    evals, t_eval = 0, 0.0
    while t_eval < seconds:
        for p, o in enumerate(opts):
            sl = Xs[p * starts:(p + 1) * starts]
            _, _, _, secs = o.eval_batch(sl, 2, nthreads=threads)
            evals += sl.shape[0]
            t_eval += secs
    return {"value": evals / t_eval, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d trajectories (%d waypoint sets) of the workload, repeated %.0f s, oracle "
                      "(reference-faithful dense formulation, -O2 -ffp-contract=off), %d threads" % (
                          Xs.shape[0], len(opts), t_eval, threads)}


def run_reference(args):
    """The reference's own CPU implementation of the path, timed on the host cores of rank 0 (the oracle port:
    the reference cannot be compiled in this image — no Eigen/ROS/nlopt.hpp)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from grad_traj_optimization_b200 import fixtures as fx
    from oracle import binding
    n_gpus = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    n_sets = 4
    # same generator, same seeds as the GPU arm's first waypoint sets; distance grid from the oracle's EDT
    from oracle import OracleMap
    om = OracleMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"])
    om.reset(), om.set_occupancy(fx.map_a_obstacles()), om.update_esdf()
    dist = om.distance()
    Xs = []
    for p in range(n_sets):
        w = fx.random_waypoints(p, M_SEG, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"])
        x0, lb, ub = fx.straight_line_x0(w, fx.PARAMS_OPTI["bos"], fx.PARAMS_OPTI["vos"], fx.PARAMS_OPTI["aos"])
        Xs.append(fx.multi_starts(x0, lb, ub, N_STARTS, seed=7 + p))
    Xs = np.concatenate(Xs)
    opts = _oracle_problems(Xs, N_STARTS)
    threads = binding.hardware_threads()

    def step():
        t = 0.0
        for p, o in enumerate(opts):
            t += o.eval_batch(Xs[p * N_STARTS:(p + 1) * N_STARTS], 2, nthreads=threads)[3]
        return t

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    secs = time.perf_counter() - t0
    ms_step = secs / args.steps * 1e3
    value = Xs.shape[0] / (ms_step * 1e-3)
    sample = "each step = %d trajectories (first %d waypoint sets x %d starts of the workload) on %d host threads" % (
        Xs.shape[0], n_sets, N_STARTS, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[2]: Map A 20x20x5m@0.1m, m=10 segments x 30 samples, step=2; bounded sample",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the on-device optimiser-loop leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
