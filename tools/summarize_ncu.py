#!/usr/bin/env python
"""Dev tool: turn Nsight Compute output brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python tools/summarize_ncu.py full     gpurun_out/prof_r1.ncu-rep profiles/r1_full.md

`launches`: per-kernel count / total / share of the `--metrics gpu__time_duration.sum` pass.
`full`:     one row per profiled launch of the `--set full` capture: duration, DRAM bytes (read+write), achieved
            DRAM GB/s, tensor-pipe %, SM busy %, occupancy, registers, shared memory (read with `ncu -i --page raw --csv`).
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_]+)(<.*?>)?\(", name)
    if not m:
        return name[:60]
    t = m.group(2) or ""
    t = re.sub(r"RowGemmMulti<|ConvLoader|DenseLoader", lambda x: {"RowGemmMulti<": "Multi<", "ConvLoader": "Conv", "DenseLoader": "Dense"}[x.group(0)], t)
    return (m.group(1) + t)[:70]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = None
    agg = OrderedDict()
    total = 0.0
    n = 0
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
                ik, iv, ig, ib = r.index("Kernel Name"), r.index("Metric Value"), r.index("Grid Size"), r.index("Block Size")
                iu = r.index("Metric Unit")
            continue
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v if r[iu] in ("us", "usecond") else v * 1e3)
        key = (short(r[ik]), r[ig], r[ib])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    with open(dst, "w") as f:
        f.write("# Launch list summary (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("source: `%s` — %d launches, %.1f us total (cold-cache, serialised: compare SHARES, not absolutes)\n\n" % (src, n, total))
        f.write("| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---:|---:|---:|---:|\n")
        for (k, g, b), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %s | %s | %d | %.1f | %.2f | %.1f%% |\n" % (k, g, b, c, t, t / c, 100 * t / total))
    print("wrote", dst)


FULL_COLS = [
    ("gpu__time_duration.sum", "dur us", 1.0),
    ("dram__bytes_read.sum", "DRAM rd MB", 1.0),
    ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
    ("lts__t_bytes.sum", "L2 MB", 1.0),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %act", 1.0),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma inst %", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__shared_mem_per_block_dynamic", "dsmem KB", 1.0),
]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik, ig, ib = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    cols = []
    for name, label, _ in FULL_COLS:
        cand = [i for i, h in enumerate(hdr) if h == name or h.endswith("." + name)]
        if not cand and "tensor" in name:
            cand = [i for i, h in enumerate(hdr) if "pipe_tensor" in h and "cycles_active" in h and h.endswith("pct_of_peak_sustained_active")]
        cols.append((label, cand[0] if cand else None))

    def conv(v, u):
        v = float(v.replace(",", "")) if v not in ("", "n/a") else float("nan")
        scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3,
                 "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
        return v * scale

    with open(dst, "w") as f:
        f.write("# `ncu --set full --clock-control none --import-source on` summary\n\nsource: `%s` (read with `ncu -i … --page raw --csv`); "
                "one row per profiled launch; ncu replays each launch with cold caches, so durations are upper bounds.\n\n" % src)
        f.write("| kernel | grid | block | " + " | ".join(l for l, _ in cols) + " | DRAM GB/s |\n")
        f.write("|---|---|---|" + "---:|" * (len(cols) + 1) + "\n")
        for r in rows[2:]:
            vals = []
            d = {}
            for l, i in cols:
                if i is None:
                    vals.append("-")
                    continue
                v = conv(r[i], units[i])
                d[l] = v
                vals.append("%.2f" % v if v < 100 else "%.0f" % v)
            gbs = (d.get("DRAM rd MB", 0) + d.get("DRAM wr MB", 0)) * 1e6 / (d.get("dur us", 1) * 1e-6) / 1e9 if d.get("dur us") else float("nan")
            f.write("| `%s` | %s | %s | %s | %.0f |\n" % (short(r[ik]), r[ig], r[ib], " | ".join(vals), gbs))
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
