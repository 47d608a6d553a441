#!/usr/bin/env python
"""Summarise ncu captures into small text files under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x.md ["title"]
    python tools/ncu_summary.py --launches gpurun_out/launches_r01.csv profiles/r01_launches.md
    python tools/ncu_summary.py --traffic f32 message=gpurun_out/a.ncu-rep gru=gpurun_out/b.ncu-rep [pack=...]
        -> updates profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch; read by bench.py)
    python tools/ncu_summary.py --sass ptgnn_b200/libptgnn_b200.so profiles/r02_sass_opcodes.md
        -> opcode histogram of the shipped library (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA, LDGSTS = cp.async)
"""
import json
import os
import re
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarise(rep, dst, title):
    hdr, units, rows = raw_page(rep)
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on; one row per captured launch)", ""]
    for r in rows:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"## {name}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"| {m} | {r[i]} | {units[i]} |")
        tensor = [(h, r[hdr.index(h)]) for h in hdr if "tensor" in h.lower() and "realtime" in h and "pct" in h]
        for h, v in tensor:
            lines.append(f"| {h} | {v} | % |")
        i_t = hdr.index("gpu__time_duration.sum")
        t = float(r[i_t]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(units[i_t], 1e-6)
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * scale.get(units[hdr.index("dram__bytes_read.sum")], 1e6)
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * scale.get(units[hdr.index("dram__bytes_write.sum")], 1e6)
        lines.append(f"| derived: DRAM traffic (read+write) | {(rd + wr) / 1e6:.1f} | MB |")
        lines.append(f"| derived: DRAM bandwidth | {(rd + wr) / t / 1e9:.0f} | GB/s |")
        lines.append("")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= i_val:
            continue
        v = float(r[i_val].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[i_unit], 1.0)
        name = r[i_name].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    lines = ["# kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)", "",
             f"source: `{src}`; {sum(a[0] for a in agg.values())} launches, {total / 1e3:.2f} ms total", "",
             "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {name} | {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f}% |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


def traffic(dtype, pairs):
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    try:
        doc = json.load(open(path))
    except (OSError, ValueError):
        doc = {}
    entry = doc.setdefault(dtype, {"kernels": {}, "source": ""})
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    srcs = []
    for pair in pairs:
        name, rep = pair.split("=", 1)
        hdr, units, rows = raw_page(rep)
        vals = []
        for r in rows:
            rd = float(r[hdr.index("dram__bytes_read.sum")]) * scale.get(units[hdr.index("dram__bytes_read.sum")], 1e6)
            wr = float(r[hdr.index("dram__bytes_write.sum")]) * scale.get(units[hdr.index("dram__bytes_write.sum")], 1e6)
            vals.append(rd + wr)
        entry["kernels"][name] = sum(vals) / len(vals)
        srcs.append(f"{name}: {os.path.basename(rep)} ({rows[0][hdr.index('Kernel Name')].split('(')[0]}, {len(vals)} launch(es))")
    entry["source"] = "ncu --set full --clock-control none, config 2; " + "; ".join(srcs)
    json.dump(doc, open(path, "w"), indent=1)
    print("wrote", path, entry)


def sass_histogram(lib, dst):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts, per_fn, fn = {}, {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and fn:
            op = m.group(1).split(".")[0]
            counts[op] = counts.get(op, 0) + 1
            if op.startswith("UTC") or op in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "SYNCS", "USETMAXREG"):
                per_fn.setdefault(fn, {}).setdefault(op, 0)
                per_fn[fn][op] += 1
    keys = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "USETMAXREG", "HMMA", "FFMA"]
    lines = [f"# SASS opcode histogram of `{lib}` (cuobjdump -sass; sm_100a)", "",
             "Blackwell-native evidence: `UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st, `UTMALDG` = TMA tile loads, `LDGSTS` = cp.async, "
             "`SYNCS` = mbarrier ops, `USETMAXREG` = setmaxnreg.  No `HMMA` (legacy mma.sync) anywhere.", "",
             "| opcode | count (whole library) |", "|---|---|"]
    for k in keys:
        lines.append(f"| {k} | {counts.get(k, 0)} |")
    lines += ["", "## tensor / TMA / TMEM opcodes per kernel (template instantiations merged)", "", "| kernel | opcodes |", "|---|---|"]
    merged = {}
    for fn, ops in per_fn.items():
        base = re.sub(r"<.*", "", fn)
        m = merged.setdefault(base, {})
        for k, v in ops.items():
            m[k] = m.get(k, 0) + v
    for fn, ops in sorted(merged.items()):
        lines.append(f"| {fn} | " + ", ".join(f"{k} {v}" for k, v in sorted(ops.items())) + " |")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "--traffic":
        traffic(sys.argv[2], sys.argv[3:])
    elif sys.argv[1] == "--sass":
        sass_histogram(sys.argv[2], sys.argv[3])
    else:
        summarise(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
