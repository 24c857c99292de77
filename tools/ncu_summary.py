#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline counters, SASS opcode mix per work unit, stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--units 655360] [--kernel-index 0] [--hot N]
"""
import argparse
import collections
import csv
import io
import subprocess


KERNEL = None  # --kernel regex: restrict both pages to one kernel of a multi-kernel report


def ncu_csv(rep, page):
    cmd = ["ncu", "-i", rep, "--page", page, "--csv"] + (["-k", "regex:" + KERNEL] if KERNEL else [])
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--units", type=float, default=655360.0, help="work units per launch (e.g. trajectory-segments)")
    ap.add_argument("--kernel-index", type=int, default=0)
    ap.add_argument("--kernel", default=None, help="regex: pick this kernel of a multi-kernel report")
    ap.add_argument("--hot", type=int, default=0, help="print the N SASS instructions with most stall samples")
    args = ap.parse_args()
    global KERNEL
    KERNEL = args.kernel

    raw = ncu_csv(args.rep, "raw")
    hdr, units, data = raw[0], raw[1], raw[2:]
    row = data[args.kernel_index]
    name_i = hdr.index("Kernel Name")
    print("kernel:", row[name_i][:100])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("  %-70s %s %s" % (k, row[i], units[i]))

    src = ncu_csv(args.rep, "source")
    blocks, cur, hdr2 = [], None, None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif r and r[0] == "Address":
            hdr2 = r
        elif cur is not None and r:
            cur.append(r)
    b = blocks[args.kernel_index]
    ix = {h: i for i, h in enumerate(hdr2)}
    byop, stall = collections.Counter(), collections.Counter()
    tot = samples = 0
    hot = []
    for r in b:
        try:
            n = int(r[ix["Instructions Executed"]])
        except ValueError:
            continue
        s = r[ix["Source"]].split()
        op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        byop[op] += n
        tot += n
        ns = int(r[ix["# Samples"]] or 0)
        samples += ns
        hot.append((ns, n, r[ix["Source"]]))
        for k in hdr2:
            if k.startswith("stall_") and "Not Issued" not in k:
                stall[k] += int(r[ix[k]] or 0)
    print("warp instructions: %d total, %.1f per unit (%d static)" % (tot, tot / args.units, len(b)))
    dp = sum(byop[o] for o in ("DADD", "DMUL", "DFMA", "DSETP"))
    print("  FP64 (DADD+DMUL+DFMA+DSETP): %.1f per unit" % (dp / args.units))
    print("  " + "  ".join("%s %.1f" % (o, c / args.units) for o, c in byop.most_common(36)))
    print("stall samples: %d" % samples)
    print("  " + "  ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / max(samples, 1)) for k, v in stall.most_common(10)))
    if args.hot:
        hot.sort(reverse=True)
        for ns, n, s in hot[:args.hot]:
            print("  %6d samples  x%-5.1f %s" % (ns, n / args.units, s))


if __name__ == "__main__":
    main()
