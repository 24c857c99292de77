#!/usr/bin/env python
"""bench.py — headline measurement of the GTOP inner loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling strong|weak]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): Map A — 20x20x5 m @0.1 m
random-pillar map — 64 waypoint sets x 1,024 multi-starts = 65,536 trajectories x 10 segments x 30 samples.
One "step" = one batched cost+gradient evaluation of the batch.  Multi-GPU (one process per GPU, torchrun): the
65,536 batch is SHARDED in contiguous blocks (`--scaling strong`, the default and what BASELINE.json / SURVEY §8e
state: 8,192 per GPU at 8 GPUs), the grid replicated, no collective on the data path (only the timing
all-reduce); `--scaling weak` keeps 65,536 per GPU, and a strong run at N > 1 adds that figure as `weak_scaling`.

`value` = evaluations/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` =
the same through the reference-facing C ABI call gtop_eval_batch_host with pinned HOST buffers (H2D of x and D2H of
cost+grad inside the timed region).  At N = 1 rank 0 also reports the other named shapes (`configs`: BASELINE
configs[1], [3], [4]), the EDT rebuilds (`edt`) and the CPU baseline.

--impl reference times the reference's own CPU implementation of the path on the host cores of rank 0: its
sources compiled against the stand-in headers of oracle/shim (oracle/_ref/libgtop_reference.so, built where
/root/reference exists and shipped with the snapshot), else the oracle port; one step = a bounded sample of the
same workload.
"""
import argparse
import json
import math
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
L2_BYTES = 126e6
WORKLOAD = ("configs[2]: Map A 20x20x5m@0.1m (200x200x50 fp64 SDF), 64 waypoint sets x 1024 multi-starts = 65,536 "
            "trajectories, m=10 segments x 30 samples, step=2 (opti_node.launch params)")


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
            try:  # fall back to one nvidia-smi sample
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


def ncu_traffic():
    """dram__bytes_read+write per launch of the dominant kernel from the committed `ncu --set full` capture: a static
    number read from profiles/, NOT measured in this run."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        d = json.load(open(p))
        return d.get("k_eval_mma_dram_bytes_per_launch"), "static: %s (not measured in this run)" % d.get("source")
    except Exception:
        return None, None


def map_a(api, fx, device):
    gm = api.SdfMap(fx.MAP_A["map_origin"], fx.MAP_A["resolution"], fx.MAP_A["map_size"], device=device)
    gm.update(fx.map_a_obstacles())
    return gm


def config3_batch(api, fx, gm, device, first_set=0):
    """64 waypoint sets x 1,024 starts on Map A: (ProblemSet, X[65536, 81], starts per problem)."""
    dist = gm.distance()
    wps = np.stack([fx.random_waypoints(first_set + p, M_SEG, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"])
                    for p in range(N_SETS)])
    ps = api.ProblemSet.from_paths(wps, api.make_params(fx.PARAMS_OPTI, samples_per_segment=N_SAMP), device=device)
    x0, (lb, ub) = ps.x0(), ps.bounds()
    X = np.concatenate([fx.multi_starts(x0[p], lb[p], ub[p], N_STARTS, seed=7 + first_set + p) for p in range(N_SETS)])
    return ps, X, N_STARTS


class DeviceTimer:
    """Device-resident evaluation of one batch, timed with CUDA events on the launching stream.  The X / grad / cost
    buffer sets rotate over more than the L2 capacity, so no step finds its inputs or its output lines in L2."""

    def __init__(self, torch, api, gm, ps, X, starts, traj_offset=0, step=2):
        self.torch, self.api, self.gm, self.ps, self.starts, self.off, self.step = torch, api, gm, ps, starts, traj_offset, step
        B, n = X.shape
        self.B, self.n = B, n
        self.nbuf = int(min(64, max(4, math.ceil(2.7 * L2_BYTES / (B * (2 * n + 1) * 8.0)))))
        self.stream = torch.cuda.Stream()
        self.Xd = [torch.from_numpy(X).cuda() for _ in range(self.nbuf)]
        self.Gd = [torch.empty((B, n), dtype=torch.float64, device="cuda") for _ in range(self.nbuf)]
        self.Cd = [torch.empty((B,), dtype=torch.float64, device="cuda") for _ in range(self.nbuf)]
        torch.cuda.synchronize()

    def launch(self, i):
        j = i % self.nbuf
        self.ps.eval_device(self.gm, self.Xd[j], self.Cd[j], self.Gd[j], step=self.step, traj_offset=self.off,
                            starts_per_problem=self.starts, stream=self.stream.cuda_stream)

    def run(self, steps, warmup, barrier=None):
        torch = self.torch
        for i in range(max(warmup, 3)):
            self.launch(i)
        (barrier or torch.cuda.synchronize)()
        l0 = self.api.lib().gtop_kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for i in range(steps):
            self.launch(i)
        e1.record(self.stream)
        (barrier or torch.cuda.synchronize)()
        return e0.elapsed_time(e1) / steps, self.api.lib().gtop_kernel_launch_count() - l0

    def working_set_mb(self):
        return self.nbuf * self.B * (2 * self.n + 1) * 8 / 1e6


def shape_record(name, ms, B, m, ns, n, peak, extra=None):
    bytes_eval = algorithmic_bytes_per_eval(m, ns, n)
    gbs = B * bytes_eval / (ms * 1e-3) / 1e9
    out = {"workload": name, "batch": B, "segments": m, "samples_per_segment": ns, "ms_per_batch": ms,
           "evals_per_s": B / (ms * 1e-3), "algorithmic_bytes_per_eval": bytes_eval,
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}}
    out.update(extra or {})
    return out


def edt_record(gm, fill, reps, peak, name):
    """Rebuild of a whole map: per-pass kernel times, rebuild of the gather copy, start of updateESDF3d to ready."""
    best = None
    for _ in range(reps):
        gm.reset(), fill(), gm.update_esdf()
        ms = gm.last_update_ms()
        best = ms if best is None or ms[5] < best[5] else best
    nv = gm.nvox
    quad, resident = gm.gather_layout()
    # 25 B/voxel: occupancy read, two int32 round trips, fp64 write; the quad copy adds 8 read + 32 written
    bytes_total = nv * (25 + (40 if quad else 0))
    return {"map": name, "grid": list(gm.grid), "voxels": nv, "ms_z": best[0], "ms_y": best[1], "ms_x": best[2],
            "ms_edt": best[3], "ms_gather_copy": best[4], "ms_update_to_evaluator_ready": best[5],
            "gather_layout": "xy-quad copy" if quad else "distance field in 8x8x4 bricks (no copy)", "resident_bytes": resident,
            "edt_roofline": {"bound": "hbm", "achieved": nv * 25 / (best[3] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": nv * 25 / (best[3] * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_voxel": 25},
            "ready_GBps_incl_gather_copy": bytes_total / (best[5] * 1e-3) / 1e9}


def other_shapes(torch, api, fx, gm, peak, steps):
    """BASELINE configs[1], [4], [3] and the EDT rebuilds (rank 0, N = 1)."""
    out, edt = [], []
    dist = gm.distance()
    obs = fx.map_a_obstacles()
    edt.append(edt_record(gm, lambda: gm.set_occupancy(obs), 5, peak, "Map A 20x20x5m@0.1m"))
    # configs[1]: 1,024 distinct 7-waypoint problems, one start each (general kernel)
    m, B = 6, 1024
    wps = np.stack([fx.random_waypoints(p, m, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"]) for p in range(B)])
    ps = api.ProblemSet.from_paths(wps, api.make_params(fx.PARAMS_OPTI), setup="device")
    ms, _ = DeviceTimer(torch, api, gm, ps, ps.x0(), 1).run(max(steps, 50), 5)
    out.append(shape_record("configs[1]: 1,024 distinct 7-waypoint problems on Map A (k_eval_traj_seg)", ms, B, m, 30, ps.n, peak))
    del ps
    # configs[4]: m = 30, 200 samples per segment, 16 waypoint sets x 1,024 starts
    m, ns, sets = 30, 200, 16
    wps = np.stack([fx.random_waypoints(p, m, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"]) for p in range(sets)])
    ps = api.ProblemSet.from_paths(wps, api.make_params(dict(fx.PARAMS_OPTI, samples_per_segment=ns)))
    x0, (lb, ub) = ps.x0(), ps.bounds()
    X = np.concatenate([fx.multi_starts(x0[p], lb[p], ub[p], N_STARTS, seed=7 + p) for p in range(sets)])
    ms, _ = DeviceTimer(torch, api, gm, ps, X, N_STARTS).run(max(4, min(steps, 20)), 3)
    out.append(shape_record("configs[4]: 30 segments x 200 samples, 16 waypoint sets x 1,024 starts on Map A (k_eval_mma)",
                            ms, X.shape[0], m, ns, ps.n, peak))
    del ps, X
    # configs[3]: 100x100x20 m @0.05 m (2000x2000x400 voxels): EDT rebuild + 4,096 distinct problems
    torch.cuda.empty_cache()
    gb = api.SdfMap(fx.MAP_B["map_origin"], fx.MAP_B["resolution"], fx.MAP_B["map_size"])
    boxes = fx.box_pillars(3000, 4, gb.grid)
    edt.append(edt_record(gb, lambda: gb.set_occupancy_boxes(boxes), 3, peak, "Map B 100x100x20m@0.05m (3,000 box pillars)"))
    m, B = 10, 4096
    free = fx.boxes_clearance_fn(boxes, fx.MAP_B["map_origin"], fx.MAP_B["resolution"], 0.5)
    wps = np.stack([fx.random_waypoints_free(p, m, free, 46.0, -45.0 + 80.0 * ((p * 0.6180339887) % 1.0), start_y_span=40.0)
                    for p in range(B)])
    ps = api.ProblemSet.from_paths(wps, api.make_params(fx.PARAMS_OPTI), setup="device")
    t = DeviceTimer(torch, api, gb, ps, ps.x0(), 1)
    ms, _ = t.run(max(steps, 50), 5)
    c = t.Cd[0].cpu().numpy()
    out.append(shape_record("configs[3]: 4,096 distinct 11-waypoint problems on Map B, 12.8 GB field in HBM in 8x8x4 bricks (k_eval_traj_seg; every step evaluates the same samples, as an optimiser loop does)",
                            ms, B, m, 30, ps.n, peak, {"finite": bool(np.isfinite(c).all()), "out_of_map_samples": ps.take_out_of_map()}))
    del t, ps, gb
    torch.cuda.empty_cache()
    return out, edt


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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    full = N_SETS * N_STARTS
    gm = map_a(api, fx, local)
    # strong: every rank builds the same global batch and takes its contiguous block; weak: its own 65,536
    ps, Xfull, starts = config3_batch(api, fx, gm, local, first_set=rank * N_SETS if args.scaling == "weak" else 0)
    per_gpu = full if args.scaling == "weak" else full // n_gpus
    lo = 0 if args.scaling == "weak" else rank * per_gpu
    X, traj_offset = Xfull[lo:lo + per_gpu], lo
    B, n = X.shape
    timer = DeviceTimer(torch, api, gm, ps, X, starts, traj_offset)
    head_nbuf, head_ws = timer.nbuf, timer.working_set_mb()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-resident arm -------------------------------------------------------------------
    sampler.active = True
    ms_step, launches = timer.run(args.steps, args.warmup, barrier)
    sampler.active = False
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
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    sampler.active = False
    # device result and host-path result of the same inputs must be identical
    same = bool(np.array_equal(Ch.numpy(), timer.Cd[0].cpu().numpy()))

    # ---- optimiser loop on the device: the shard through 100 lock-step LD_MMA evaluations; device-
    # resident and through the host entry point (x in, x + min f out: what crosses PCIe when the loop lives on the GPU)
    solve = None
    if not args.no_solve:
        try:
            work = timer.Xd[0].clone()
            minf = torch.empty((B,), dtype=torch.float64, device="cuda")
            ps.solve_device(gm, work, minf, step=2, maxeval=5, traj_offset=traj_offset, starts_per_problem=starts,
                            stream=timer.stream.cuda_stream)  # warm-up
            work.copy_(timer.Xd[0])
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(timer.stream)
            ps.solve_device(gm, work, minf, step=2, maxeval=100, traj_offset=traj_offset, starts_per_problem=starts,
                            stream=timer.stream.cuda_stream)
            s1.record(timer.stream)
            barrier()
            solve_ms = s0.elapsed_time(s1)
            Xs, Ms = Xh.clone().pin_memory(), Ch.clone().pin_memory()  # pinned: x in, x + min f out, solved in place
            ps.solve_host(gm, Xs, step=2, maxeval=2, traj_offset=traj_offset, starts_per_problem=starts, minf=Ms)  # warm-up
            Xs.copy_(Xh)
            barrier()
            t0 = time.perf_counter()
            ps.solve_host(gm, Xs, step=2, maxeval=100, traj_offset=traj_offset, starts_per_problem=starts, minf=Ms)
            solve_host_ms = (time.perf_counter() - t0) * 1e3
            solve_ms, solve_host_ms = max_over_ranks(solve_ms, solve_host_ms)
            solve = {"maxeval": 100, "ms": solve_ms, "solves_per_s": B * n_gpus / solve_ms * 1e3,
                     "e2e_ms": solve_host_ms, "e2e_solves_per_s": B * n_gpus / solve_host_ms * 1e3,
                     "e2e_h2d_bytes": B * n * 8, "e2e_d2h_bytes": B * n * 8 + B * 8,
                     "median_cost_start": float(torch.median(timer.Cd[0]).item()), "median_cost_end": float(torch.median(minf).item())}
        except Exception as exc:  # the headline measurement above does not depend on this leg
            solve = {"error": str(exc)}

    # ---- weak-scaling figure beside a strong run (65,536 per GPU, the round-1 mode) -----------------------------
    weak = None
    if args.scaling == "strong" and world > 1:
        del timer
        torch.cuda.empty_cache()
        timer = DeviceTimer(torch, api, gm, ps, Xfull, starts, 0)
        wms, _ = timer.run(max(5, min(args.steps, 40)), 3, barrier)
        (wms,) = max_over_ranks(wms)
        weak = {"per_gpu_batch": full, "global_batch": full * n_gpus, "ms_per_step": wms, "value": full * n_gpus / (wms * 1e-3)}

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

    ms_step, e2e_ms = max_over_ranks(ms_step, e2e_ms)
    sampler.stop_flag = True
    l2_note = ("%d rotating X/grad/cost buffer sets per GPU (%.0f MB > 126 MB L2); the 16 MB SDF grid is L2-resident by "
               "construction" % (head_nbuf, head_ws))

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        total_evals = B * n_gpus
        value = total_evals / (ms_step * 1e-3)
        bytes_eval = algorithmic_bytes_per_eval(M_SEG, N_SAMP, n)
        achieved = B * bytes_eval / (ms_step * 1e-3) / 1e9  # per-GPU kernel: one launch = B evaluations
        traffic, traffic_src = ncu_traffic()
        configs = edt = None
        if world == 1 and not args.no_configs:
            try:
                timer = None
                torch.cuda.empty_cache()
                configs, edt = other_shapes(torch, api, fx, gm, peak, min(args.steps, 100))
            except Exception as exc:
                configs = {"error": str(exc)}
        cpu = cpu_baseline(Xfull[:4096], 12.0) if (world == 1 and not args.no_cpu_baseline) else None
        if fp64 and "dfma_peak_tflops" in fp64:
            # executed FP64 work of k_eval_mma per trajectory-segment (ncu, profiles/r2_k_eval_mma_ncu.txt): 99.7
            # DADD/DMUL + 32.4 DFMA warp instructions (32 lanes) + 6.0 DMMA (512 flops each)
            flops_unit = (52.8 + 46.9) * 32 + 32.4 * 64 + 6.0 * 512
            tf = flops_unit * M_SEG * B / (ms_step * 1e-3) / 1e12
            fp64.update({"kernel_executed_tflops": tf, "frac_of_dfma_peak": tf / fp64["dfma_peak_tflops"]})
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": total_evals, "segments": M_SEG,
                       "samples_per_segment": N_SAMP,
                       "parallelism": "batch %s across %d GPU(s) in contiguous blocks, SDF replicated, no collective" % (
                           "sharded" if args.scaling == "strong" else "replicated per GPU (weak)", n_gpus),
                       "l2": l2_note,
                       "out_of_map_samples_per_step": oob // max(1, args.steps + max(args.warmup, 3))},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind, "kernel": "k_eval_mma",
                         "algorithmic_bytes_per_eval": bytes_eval, "kernel_ms": ms_step},
            "cpu_baseline": cpu,
            "e2e": {"value": total_evals / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * n * 8,
                    "d2h_bytes_per_step": B * n * 8 + B * 8, "steps": e2e_steps, "host_equals_device": same},
            "gpu_launches": int(launches),
            "weak_scaling": weak,
            "configs": configs,
            "edt": edt,
            "solve_on_device": solve,
            "fp64": fp64,
            "clocks": sampler.summary(),
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation of the path (its sources against oracle/shim), else the oracle port
# ------------------------------------------------------------------------------------------------------------------
class CpuPath:
    """costFunc of the reference on all host threads: one GradTrajOptimizer per thread and waypoint set (the class is
    not re-entrant, distinct objects are independent once set up; set-up runs serially — qp_generator.cpp keeps
    globals).  Falls back to the oracle port when the reference build is not there."""

    def __init__(self, n_sets):
        from grad_traj_optimization_b200 import fixtures as fx
        from oracle import binding
        from oracle import ref_binding as rb
        self.threads = binding.hardware_threads()
        self.kind = "reference" if rb.available() else "port"
        self.n_sets = n_sets
        cfg = dict(map_size=fx.MAP_A["map_size"], map_origin=fx.MAP_A["map_origin"], resolution=fx.MAP_A["resolution"])
        if self.kind == "reference":
            rm = rb.ReferenceMap(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
            rm.reset(), rm.set_occupancy(fx.map_a_obstacles())
            t0 = time.perf_counter()
            rm.update_esdf()
            self.edt_ms = (time.perf_counter() - t0) * 1e3
            dist = rm.distance()
            self.workers = []
            for _ in range(self.threads):
                row = []
                for p in range(n_sets):
                    ro = rb.ReferenceOptimizer(fx.PARAMS_OPTI)
                    ro.init_map(cfg["map_size"], cfg["map_origin"], cfg["resolution"])
                    ro.set_distance(dist)
                    ro.set_path(fx.random_waypoints(p, M_SEG, dist, cfg["map_origin"], cfg["resolution"]))
                    row.append(ro)
                self.workers.append(row)
            self.what = ("the reference's src/{sdf_map,qp_generator,grad_traj_optimizer}.cpp compiled -O3 against the "
                         "stand-in headers of oracle/shim (its Eigen / ROS / nlopt.hpp are not in this image)")
        else:
            from oracle import OracleMap, OracleOptimizer, OracleParams
            om = OracleMap(cfg["map_origin"], cfg["resolution"], cfg["map_size"])
            om.reset(), om.set_occupancy(fx.map_a_obstacles())
            t0 = time.perf_counter()
            om.update_esdf()
            self.edt_ms = (time.perf_counter() - t0) * 1e3
            dist = om.distance()
            prm = {"w_smooth": fx.PARAMS_OPTI["ws"], "w_collision": fx.PARAMS_OPTI["wc"], "samples_per_segment": N_SAMP,
                   **{k: v for k, v in fx.PARAMS_OPTI.items() if k not in ("ws", "wc", "alg", "dt")}}
            self.opts = []
            for p in range(n_sets):
                o = OracleOptimizer(OracleParams.from_dict(prm), om)
                o.set_path(fx.random_waypoints(p, M_SEG, dist, cfg["map_origin"], cfg["resolution"]))
                self.opts.append(o)
            self.what = "oracle port of the reference objective (-O2 -ffp-contract=off; the reference build is absent)"

    def evaluate(self, X):
        """X: rows in workload order (set p = rows p*1024 ..).  Returns wall seconds of evaluating all of them."""
        t0 = time.perf_counter()
        if self.kind == "port":
            for p in range(self.n_sets):
                sl = X[p * N_STARTS:(p + 1) * N_STARTS]
                if len(sl):
                    self.opts[p].eval_batch(sl, 2, nthreads=self.threads)
            return time.perf_counter() - t0
        by_worker = {}  # worker (thread) index -> [(waypoint set, rows)]
        for p in range(self.n_sets):
            sl = X[p * N_STARTS:(p + 1) * N_STARTS]
            for w, part in enumerate(np.array_split(sl, self.threads) if len(sl) else []):
                if len(part):
                    by_worker.setdefault(w, []).append((p, part))
        th = [threading.Thread(target=lambda w=w, lst=lst: [self.workers[w][p].eval_batch(part, 2) for p, part in lst])
              for w, lst in by_worker.items()]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0


def cpu_baseline(Xs, seconds):
    """The CPU implementation on all host cores, on the first 4,096 trajectories of the workload (4 waypoint sets),
    repeated for ~`seconds`.  A reported baseline, not the optimisation target."""
    cpu = CpuPath(max(1, Xs.shape[0] // N_STARTS))
    probe = Xs[:max(cpu.threads * 4, 64)]
    cpu.evaluate(probe)
    rate = len(probe) / cpu.evaluate(probe)
    take = int(min(Xs.shape[0], max(cpu.threads * 8, rate * 2.0)))  # ~2 s per pass
    evals, t = 0, 0.0
    while t < seconds:
        t += cpu.evaluate(Xs[:take])
        evals += take
    return {"value": evals / t, "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind, "evals_per_s_per_core": evals / t / cpu.threads,
            "edt_ms_1thread": cpu.edt_ms, "edt_map": "Map A 200x200x50, updateESDF3d, one thread",
            "sample": "first %d trajectories of the workload, repeated %.0f s on %d threads: %s" % (take, t, cpu.threads, cpu.what)}


def run_reference(args):
    """The reference's own CPU implementation of the path, timed on the host cores of rank 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from grad_traj_optimization_b200 import fixtures as fx
    n_gpus = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    n_sets = 4
    cpu = CpuPath(n_sets)
    dist = (cpu.workers[0][0].distance() if cpu.kind == "reference" else cpu.opts[0].map.distance())
    Xs = []
    for p in range(n_sets):  # same generator, same seeds as the GPU arm's first waypoint sets
        w = fx.random_waypoints(p, M_SEG, dist, fx.MAP_A["map_origin"], fx.MAP_A["resolution"])
        x0, lb, ub = fx.straight_line_x0(w, fx.PARAMS_OPTI["bos"], fx.PARAMS_OPTI["vos"], fx.PARAMS_OPTI["aos"])
        Xs.append(fx.multi_starts(x0, lb, ub, N_STARTS, seed=7 + p))
    Xs = np.concatenate(Xs)
    # a step = a bounded sample of the 65,536 batch, sized so that warm-up + K steps take about two minutes
    probe = Xs[:max(cpu.threads * 4, 64)]
    cpu.evaluate(probe)
    rate = len(probe) / cpu.evaluate(probe)
    take = int(min(Xs.shape[0], max(cpu.threads, rate * 120.0 / (args.steps + max(args.warmup, 1)))))
    for _ in range(max(args.warmup, 1)):
        cpu.evaluate(Xs[:take])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.evaluate(Xs[:take])
    secs = time.perf_counter() - t0
    ms_step = secs / args.steps * 1e3
    value = take / (ms_step * 1e-3)
    sample = "each step = the first %d trajectories of the 65,536 batch on %d host threads: %s" % (take, cpu.threads, cpu.what)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "per_step_batch": take},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind, "sample": sample,
                         "evals_per_s_per_core": value / cpu.threads, "edt_ms_1thread": cpu.edt_ms},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--scaling", choices=["strong", "weak"], default="strong")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the on-device optimiser-loop leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the other named shapes and the EDT rebuilds (N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
