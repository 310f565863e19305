#!/usr/bin/env python
"""Summaries of ncu captures for profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/x.csv "command line" > profiles/x_summary.md
    python tools/ncu_summary.py full gpurun_out/x.ncu-rep profiles/x.json > profiles/x_summary.md
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", name)[:110]


def launches(path, cmd):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1; a[1] += us
        n += 1
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "at::" not in k and "cutlass" not in k and "cublas" not in k and "nvjet" not in k
               and "gemv" not in k and "memcpy" not in k.lower() and "elementwise" not in k)
    print("Command (after the same command exited 0 without ncu): `ncu --metrics gpu__time_duration.sum --clock-control none --csv %s`\n" % cmd)
    print("%d launches, %.1f ms of kernel time; per-launch times are cold-cache and serialised: read SHARES.\n" % (n, tot / 1e3))
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, c, t, 100 * t / tot, t / c))
    print("\nlibekl_b200 kernels: %.1f%% of GPU time in this window." % (100 * ours / tot))


def full(rep, out_json):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    ki = hdr.index("Kernel Name")
    out = []
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        d = {"kernel": short(r[ki])}
        for w, i in idx:
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            d[w] = v * mult.get(units[i], 1) if "bytes" in w else v
            if w == "gpu__time_duration.sum":
                d[w] = v / 1e3 if units[i].startswith("n") else v
        out.append(d)
    json.dump(out, open(out_json, "w"), indent=1)
    print("| kernel | us | DRAM read MB | DRAM write MB | DRAM %% | tensor pipe %% | L2 hit %% | L2->SM GB | regs | grid |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for d in out:
        g = lambda k, s=1.0: ("%.1f" % (d[k] / s)) if k in d else "-"
        print("| `%s` | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
            d["kernel"], g("gpu__time_duration.sum"), g("dram__bytes_read.sum", 1e6), g("dram__bytes_write.sum", 1e6),
            g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            g("lts__t_sector_hit_rate.pct"), g("l1tex__m_xbar2l1tex_read_bytes.sum", 1e9), g("launch__registers_per_thread"), g("launch__grid_size")))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3])
