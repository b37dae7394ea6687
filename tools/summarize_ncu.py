#!/usr/bin/env python
"""Turns ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py r01
reads gpurun_out/prof_<tag>_*.ncu-rep (ncu --set full, one launch each) and gpurun_out/launches_<tag>.csv
(ncu --metrics gpu__time_duration.sum) and writes profiles/<tag>_<kernel>.csv (selected raw metrics),
profiles/<tag>_launches.csv (per-launch device times) and profiles/<tag>_summary.md.
"""
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
]
STALLS = "smsp__average_warps_issue_stalled_"

def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


# bench.py kernel name of each capture (bench.py reads <tag>_traffic.json for roofline.traffic)
BENCH_NAME = {"blur_dog_o0": "blur_dog_down_o0", "pyramid_o0": "pyramid_o0", "pyramid_rest": "pyramid_rest",
              "down_chain": "down_chain", "find_points": "find_points", "orient_desc": "orient_desc", "match_tc": "match_tc"}
traffic = {}
PROF.mkdir(exist_ok=True)
lines = [f"# ncu summaries, tag {tag}", "",
         "Captured on a B200 with `ncu --set full --clock-control none --import-source on` (one launch per kernel,",
         "a 1080p frame of the bench workload; cold caches, serialised) and",
         "`ncu --metrics gpu__time_duration.sum --clock-control none` (launch list).  Numbers taken under a profiler are",
         "not bench values; they explain them.", ""]
for rep in sorted(OUT.glob(f"prof_{tag}_*.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = rep.stem.replace(f"prof_{tag}_", "")
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name
    sel = [(k, vals[hdr.index(k)], units[hdr.index(k)]) for k in KEYS if k in hdr]
    stalls = sorted(((float(vals[i]), h[len(STALLS):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
                     if h.startswith(STALLS) and vals[i]), reverse=True)[:6]
    with open(PROF / f"{tag}_{name}.csv", "w", newline="") as fp:
        wr = csv.writer(fp)
        wr.writerow(["metric", "value", "unit"])
        wr.writerows(sel)
        wr.writerows((STALLS + s + "_per_issue_active.ratio", f"{v:.4f}", "") for v, s in stalls)
    d = {k: v for k, v, _ in sel}
    ud = {k: u for k, _, u in sel}
    if "dram__bytes_read.sum" in d:
        traffic[BENCH_NAME.get(name, name)] = {
            "dram_bytes": to_bytes(d["dram__bytes_read.sum"], ud["dram__bytes_read.sum"]) +
                          to_bytes(d["dram__bytes_write.sum"], ud["dram__bytes_write.sum"]),
            "dram_bytes_read": to_bytes(d["dram__bytes_read.sum"], ud["dram__bytes_read.sum"]),
            "dram_bytes_write": to_bytes(d["dram__bytes_write.sum"], ud["dram__bytes_write.sum"]),
            "capture": rep.name, "note": "one launch under ncu --set full (cache control: flush before the launch)"}
    dur_us = float(d.get("gpu__time_duration.sum", "nan").replace(",", ""))
    unit = dict((k, u) for k, _, u in sel).get("gpu__time_duration.sum", "")
    if unit == "ns":
        dur_us /= 1000.0
    rd, wrb = d.get("dram__bytes_read.sum", "?"), d.get("dram__bytes_write.sum", "?")
    lines += [f"## {name}", "", f"`{kname[:110]}`", "",
              f"* duration {dur_us:.2f} us; DRAM read {traffic[BENCH_NAME.get(name, name)]['dram_bytes_read'] / 1e6:.2f} MB / "
              f"write {traffic[BENCH_NAME.get(name, name)]['dram_bytes_write'] / 1e6:.2f} MB; "
              f"registers {d.get('launch__registers_per_thread')}, grid {d.get('launch__grid_size')} x {d.get('launch__block_size')}",
              f"* issue active {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active')} %, FMA pipe "
              f"{d.get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active')} %, ALU pipe "
              f"{d.get('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active')} %, tensor pipe "
              f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} %, warps/scheduler "
              f"{d.get('smsp__warps_active.avg.per_cycle_active')} (eligible {d.get('smsp__warps_eligible.avg.per_cycle_active')})",
              "* top stalls per issue: " + ", ".join(f"{s} {v:.2f}" for v, s in stalls), ""]
lf = OUT / f"launches_{tag}.csv"
if lf.exists():
    rows = list(csv.reader(lf.read_text().splitlines()))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    with open(PROF / f"{tag}_launches.csv", "w", newline="") as fp:
        wr = csv.writer(fp)
        wr.writerow(["kernel", "grid", "block", "gpu__time_duration_ns"])
        data = rows[start + 1:]
        for r in data:
            wr.writerow([r[ik].split("(")[0], r[ig], r[ib], r[iv]])
    per_frame = 5 if any("k_pyramid" in r[ik] for r in data) else 7   # r02: pyramid_o0, down_chain, pyramid_rest, find_points, orient_desc
    frame = data[-per_frame:]
    tot = sum(float(r[iv].replace(",", "")) for r in frame)
    lines += ["## launch list (last frame of the capture)", "", "| kernel | grid | ns | share |", "|---|---|---|---|"]
    for r in frame:
        ns = float(r[iv].replace(",", ""))
        lines.append(f"| {r[ik].split('(')[0]} | {r[ig]} | {ns:.0f} | {100*ns/tot:.1f} % |")
    lines += ["", f"sum {tot/1000:.1f} us per 1080p frame (serialised, cold cache)", ""]
bf = OUT / f"launches_bench_{tag}.csv"
if bf.exists():
    rows = list(csv.reader(bf.read_text().splitlines()))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    data = rows[start + 1:]
    with open(PROF / f"{tag}_bench_launches.csv", "w", newline="") as fp:
        wr = csv.writer(fp)
        wr.writerow(["kernel", "grid", "gpu__time_duration_ns"])
        for r in data:
            wr.writerow([r[ik].split("(")[0], r[ig], r[iv]])
    agg = {}
    for r in data:
        key = (r[ik].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", ""), r[ig])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    lines += [f"## launch list of `python bench.py --steps 2 --warmup 3 --frames-per-step 64` (first {len(data)} launches)", "",
              "| kernel | grid | launches | mean ns | share of the step |", "|---|---|---|---|---|"]
    for (k, g), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k} | {g} | {n} | {t / n:.0f} | {100 * t / tot:.1f} % |")
    lines.append("")
import json
(PROF / f"{tag}_traffic.json").write_text(json.dumps(traffic, indent=1))
(PROF / f"{tag}_summary.md").write_text("\n".join(lines))
print("\n".join(lines))
