#!/usr/bin/env python
"""Per-CUDA-source-line totals of an .ncu-rep captured with --import-source on (read here, no GPU):
warp instructions executed and stall samples per line, heaviest first.   python tools/ncu_lines.py REP [--top 40]"""
import argparse
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--by", default="samples", choices=("samples", "inst"))
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            i_s, i_n = hdr.index("# Samples"), hdr.index("Instructions Executed")
            try:
                lines.append((int(r[i_s] or 0), int(r[i_n] or 0), fpath, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    ts, tn = sum(l[0] for l in lines), sum(l[1] for l in lines)
    lines.sort(key=lambda l: -(l[0] if a.by == "samples" else l[1]))
    print("total samples %d, warp instructions %d" % (ts, tn))
    for s, n, f, ln, src in lines[:a.top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * s / max(ts, 1), 100.0 * n / max(tn, 1), f, ln, src[:110]))


if __name__ == "__main__":
    main()
