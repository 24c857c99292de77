#!/usr/bin/env python
"""Turns gpurun_out/ ncu artefacts into the committed summaries under profiles/ (round-tagged).

    python tools/make_profiles.py r1 gpurun_out/r1_launches.csv gpurun_out/r1_k_eval_traj.ncu-rep [bench.json]
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
    bench = sys.argv[4] if len(sys.argv) > 4 else None
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)

    # ---- launch list: per-kernel count / total / share (cold-cache, serialised: compare SHARES) ----
    rows = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    seq = []
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)  # -> us
        name = r[ik].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        seq.append((name, v))
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(out_dir, "%s_launches.md" % tag), "w") as f:
        f.write("# %s — ncu launch list of `python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs`\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: the SHARE "
                "of each kernel is meaningful, not the absolute time).  %d launches, %.1f us total.\n\n" % (len(seq), total))
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.2f | %.1f%% |\n" % (name, n, t, t / n, 100 * t / total))
        f.write("\nLaunch order (first 40): " + ", ".join("%s %.0fus" % (n.replace("gtop::", ""), t) for n, t in seq[:40]) + "\n")

    # ---- full capture of the dominant kernel ----
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "--hot", "20"],
                          capture_output=True, text=True).stdout
    kname = "k_eval_mma" if "mma" in os.path.basename(rep) else "k_eval_traj"
    with open(os.path.join(out_dir, "%s_%s_ncu.txt" % (tag, kname)), "w") as f:
        f.write("# %s — ncu --set full --clock-control none --import-source on -k regex:%s (one launch of the config-3 "
                "batch: 65,536 trajectories x 10 segments x 30 samples; 'unit' = one trajectory-segment = 32 sample slots "
                "= one warp iteration)\n" % (tag, kname))
        f.write(summ)
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                                                      text=True).stdout)))
    h, u, d = raw[0], raw[1], raw[2]

    def get(k):
        v = float(d[h.index(k)].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[h.index(k)], 1)

    traffic = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
    info = {"%s_dram_bytes_per_launch" % kname: traffic, "source": "profiles/%s_%s_ncu.txt" % (tag, kname),
            "kernel_time_us_under_ncu": get("gpu__time_duration.sum") if u[h.index("gpu__time_duration.sum")] == "byte" else
            float(d[h.index("gpu__time_duration.sum")].replace(",", ""))}
    json.dump(info, open(os.path.join(out_dir, "roofline_traffic.json"), "w"), indent=1)
    if bench:
        import shutil
        shutil.copy(bench, os.path.join(out_dir, "%s_bench_n1.json" % tag))
    print(open(os.path.join(out_dir, "%s_launches.md" % tag)).read())
    print(json.dumps(info))


if __name__ == "__main__":
    main()
