#!/usr/bin/env python
"""Where does gtop_eval_batch_host's time go?  The same chunked three-stream pipeline as the C entry point, rebuilt from
torch copies with the kernel replaced by nothing / by a device-side copy, one synchronous call at a time (GPU box).
   python tools/e2e_pipeline_probe.py"""
import json, time
import torch

B, n = 65536, 81
hin = torch.empty((B, n), dtype=torch.float64).pin_memory()
hg = torch.empty((B, n), dtype=torch.float64).pin_memory()
hc = torch.empty((B,), dtype=torch.float64).pin_memory()
din, dg, dc = torch.empty((B, n), dtype=torch.float64, device="cuda"), torch.empty((B, n), dtype=torch.float64, device="cuda"), torch.empty((B,), dtype=torch.float64, device="cuda")
s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def call(sizes, kernel, cost_per_chunk):
    b0 = 0
    for bn in sizes:
        sl = slice(b0, b0 + bn)
        with torch.cuda.stream(s_in):
            din[sl].copy_(hin[sl], non_blocking=True)
            e_in = torch.cuda.Event(); e_in.record(s_in)
        s_k.wait_event(e_in)
        with torch.cuda.stream(s_k):
            if kernel:
                dg[sl].copy_(din[sl])          # stands in for the evaluation kernel (reads x, writes the gradient)
            e_k = torch.cuda.Event(); e_k.record(s_k)
        s_out.wait_event(e_k)
        with torch.cuda.stream(s_out):
            hg[sl].copy_(dg[sl], non_blocking=True)
            if cost_per_chunk:
                hc[sl].copy_(dc[sl], non_blocking=True)
        b0 += bn
    if not cost_per_chunk:
        with torch.cuda.stream(s_out):
            hc.copy_(dc, non_blocking=True)
    torch.cuda.synchronize()


def timeit(sizes, kernel, cost_per_chunk, reps=30):
    for _ in range(3):
        call(sizes, kernel, cost_per_chunk)
    t0 = time.perf_counter()
    for _ in range(reps):
        call(sizes, kernel, cost_per_chunk)
    return (time.perf_counter() - t0) / reps * 1e3


out = {}
for nch in (1, 2, 4, 8, 16, 32):
    sizes = [B // nch] * nch
    out["uniform_%d" % nch] = {"copies_only": timeit(sizes, False, True), "with_kernel_standin": timeit(sizes, True, True),
                               "cost_once": timeit(sizes, True, False)}
ramp = [8192] * 7 + [4096, 2048, 1024, 512, 512]
assert sum(ramp) == B
out["ramp_down_12"] = {"copies_only": timeit(ramp, False, True), "with_kernel_standin": timeit(ramp, True, True), "cost_once": timeit(ramp, True, False)}
# the two directions alone, one big copy each
def one(direction):
    t0 = time.perf_counter()
    for _ in range(20):
        if direction == "h2d":
            din.copy_(hin, non_blocking=True)
        else:
            hg.copy_(dg, non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 20 * 1e3
out["h2d_alone_ms"], out["d2h_alone_ms"] = one("h2d"), one("d2h")
print(json.dumps(out))
