#!/usr/bin/env python3
"""Copy what is judged from gpurun_out/ (scratch) into profiles/ (tracked).

    python scripts/collect_profiles.py <tag>          e.g. r01d

  * gpurun_out/launches_<W>.csv (ncu --metrics gpu__time_duration.sum launch list of `bench.py`)  ->
        profiles/<tag>_launches_<W>.csv        per-kernel-name totals and SHARE of the GPU time of the command
        profiles/<tag>_launches_<W>_raw.csv.gz the raw list
  * gpurun_out/prof_trace_<W>.ncu-rep (ncu --set full)  ->  profiles/<tag>_k_trace_<W>.csv (scripts/ncu_summary.py) and the
        per-launch DRAM traffic of k_trace<closest> into profiles/roofline_traffic.json (read by bench.py)
  * gpurun_out/bench_*.json  ->  profiles/<tag>_bench.jsonl
"""
import csv
import glob
import gzip
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def short(name):
    m = re.match(r"(?:void )?([\w:]+(?:<[^(]*>)?)\(", name)
    s = m.group(1) if m else name
    return re.sub(r"at::native::|at::", "", s)[:80]


def launches(tag, path):
    w = re.search(r"launches_(\w+)\.csv", path).group(1)
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    tot, per = 0.0, {}
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        k = short(r["Kernel Name"])
        c = per.setdefault(k, [0, 0.0])
        c[0] += 1; c[1] += ns; tot += ns
    with open(os.path.join(PROF, f"{tag}_launches_{w}.csv"), "w", newline="") as f:
        o = csv.writer(f)
        o.writerow(["kernel", "launches", "total_ms", "mean_ms", "share_of_gpu_time"])
        for k, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            o.writerow([k, n, f"{ns / 1e6:.3f}", f"{ns / 1e6 / n:.4f}", f"{ns / tot:.4f}"])
    with gzip.open(os.path.join(PROF, f"{tag}_launches_{w}_raw.csv.gz"), "wt") as f:
        f.writelines(lines)
    print(open(os.path.join(PROF, f"{tag}_launches_{w}.csv")).read())


def traffic(tag, rep):
    w = re.search(r"prof_trace_(\w+)\.ncu-rep", rep).group(1)
    out = os.path.join(PROF, f"{tag}_k_trace_{w}.csv")
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, out], stdout=subprocess.DEVNULL, check=True)
    rows = {r[0]: r for r in csv.reader(open(out))}
    names = rows["Kernel Name"][2:]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd, wr = rows["dram__bytes_read.sum"], rows["dram__bytes_write.sum"]
    vals = [float(rd[2 + i]) * scale[rd[1]] + float(wr[2 + i]) * scale[wr[1]] for i, n in enumerate(names) if "k_trace<0" in n]
    ms = [float(rows["gpu__time_duration.sum"][2 + i]) for i, n in enumerate(names) if "k_trace<0" in n]
    tpath = os.path.join(PROF, "roofline_traffic.json")
    d = json.load(open(tpath)) if os.path.exists(tpath) else {}
    def mean_of(metric):
        r = rows.get(metric)
        xs = [float(r[2 + i]) for i, n in enumerate(names) if "k_trace<0" in n] if r else []
        return sum(xs) / len(xs) if xs else None
    keep_shade = d.get(w, {}).get("shade")
    d[w] = {"k_trace_closest_dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals), "ncu_ms_per_launch": sum(ms) / len(ms),
            # what bounds the kernel (the algorithmic-byte figure of bench.py is a model; these are the counters)
            "issue_active_pct": mean_of("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "active_lanes_per_instruction": mean_of("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "fma_pipe_active_pct": mean_of("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_active_pct": mean_of("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": mean_of("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_throughput_pct": mean_of("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1_hit_pct": mean_of("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": mean_of("lts__t_sector_hit_rate.pct"),
            "source": f"profiles/{tag}_k_trace_{w}.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, closest-hit launches of bounces 0 and 1 of one pass)"}
    if keep_shade:
        d[w]["shade"] = keep_shade
    json.dump(d, open(tpath, "w"), indent=1)
    print(w, d[w])


def shade_record(tag, w, out):
    """k_shade counters (mean over the captured launches: camera rays and the first bounce) -> roofline_traffic.json[w]["shade"]"""
    rows = {r[0]: r for r in csv.reader(open(out))}
    n = len(rows["Kernel Name"]) - 2
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd, wr = rows["dram__bytes_read.sum"], rows["dram__bytes_write.sum"]
    mean = lambda m: (sum(float(rows[m][2 + i]) for i in range(n)) / n) if m in rows else None
    tpath = os.path.join(PROF, "roofline_traffic.json")
    d = json.load(open(tpath)) if os.path.exists(tpath) else {}
    d.setdefault(w, {})["shade"] = {
        "dram_bytes_per_launch": sum(float(rd[2 + i]) * scale[rd[1]] + float(wr[2 + i]) * scale[wr[1]] for i in range(n)) / n, "launches_captured": n,
        "ncu_ms_per_launch": mean("gpu__time_duration.sum"),
        "issue_active_pct": mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "active_lanes_per_instruction": mean("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "fma_pipe_active_pct": mean("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_active_pct": mean("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": mean("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l2_throughput_pct": mean("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "l1_hit_pct": mean("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": mean("lts__t_sector_hit_rate.pct"),
        "source": f"profiles/{tag}_k_shade_{w}.csv (ncu --set full, k_shade launches of bounces 0 and 1 of one pass)"}
    json.dump(d, open(tpath, "w"), indent=1)
    print(w, "shade", d[w]["shade"])


def main():
    tag = sys.argv[1]
    for p in sorted(glob.glob(os.path.join(OUT, "launches_*.csv"))):
        launches(tag, p)
    for p in sorted(glob.glob(os.path.join(OUT, "prof_trace_*.ncu-rep"))):
        traffic(tag, p)
    for p in sorted(glob.glob(os.path.join(OUT, "prof_shade_*.ncu-rep"))):
        w = re.search(r"prof_shade_(\w+)\.ncu-rep", p).group(1)
        out = os.path.join(PROF, f"{tag}_k_shade_{w}.csv")
        subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), p, out], stdout=subprocess.DEVNULL, check=True)
        shade_record(tag, w, out)
    with open(os.path.join(PROF, f"{tag}_bench.jsonl"), "w") as f:
        for p in sorted(glob.glob(os.path.join(OUT, "bench_*.json"))):
            for l in open(p):
                if l.startswith("{"):
                    f.write(l)


if __name__ == "__main__":
    main()
