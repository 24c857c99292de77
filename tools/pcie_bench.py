#!/usr/bin/env python
"""Host <-> device copy ceiling of the end-to-end path (GPU box): every rank copies the config-3 shard sizes both ways
at the same time (pinned host memory, one stream per direction, plain cudaMemcpyAsync), all ranks together.

    python tools/pcie_bench.py                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_bench.py

Prints one JSON line (rank 0): per-GPU and aggregate GB/s each way, for the full 65,536 batch per GPU (weak) and for the
65,536 / N shard (strong) — the fabric ceiling of bench.py's `e2e` figure."""
import json
import os
import time

import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(rows, reps=20):
    n = 81
    hin = torch.empty((rows, n), dtype=torch.float64).pin_memory()
    hout = torch.empty((rows, n + 1), dtype=torch.float64).pin_memory()
    din, dout = torch.empty_like(hin, device="cuda"), torch.empty_like(hout, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for timed in (False, True):
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps if timed else 3):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s_in):
                        din.copy_(hin, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s_out):
                        hout.copy_(dout, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode + "_ms"] = float(t[0]) * 1e3
    gb_in, gb_out = hin.numel() * 8 / 1e9, hout.numel() * 8 / 1e9
    res.update({"rows_per_gpu": rows, "h2d_GBps_per_gpu_alone": gb_in / (res["h2d_ms"] * 1e-3), "d2h_GBps_per_gpu_alone": gb_out / (res["d2h_ms"] * 1e-3),
                "duplex_GBps_per_gpu_each_way": gb_in / (res["both_ms"] * 1e-3),
                "duplex_GBps_aggregate_both_ways": world * (gb_in + gb_out) / (res["both_ms"] * 1e-3),
                "e2e_ceiling_evals_per_s": world * rows / (res["both_ms"] * 1e-3)})
    return res


out = {"n_gpus": world, "weak_65536_per_gpu": run(65536), "strong_shard": run(65536 // world)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
