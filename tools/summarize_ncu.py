"""Turns gpurun_out ncu artefacts into the small tracked summaries under profiles/.
usage: summarize_ncu.py launches <launches.csv> <out.md>   |   summarize_ncu.py full <file.ncu-rep> <out.md>
       summarize_ncu.py traffic <out.json> <workload>=<file.ncu-rep> ...   (what bench.py reads for roofline.traffic / secondary.ncu)"""
import collections, csv, json, re, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
           "sm__cycles_elapsed.avg.per_second"]

def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0].replace("void ", ""), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none   (source: {src}; cold-cache serialised launches: compare SHARES)\n\n")
        f.write("| kernel | launches | avg us | share |\n|---|---:|---:|---:|\n")
        for k, v in agg.items():
            f.write(f"| {k} | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/tot*100:.1f}% |\n")

def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"ncu --set full --clock-control none --import-source on   (source: {src})\n\n")
        for r in rows[2:]:
            f.write(f"### {r[h.index('Kernel Name')].split('(')[0]}  grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in METRICS:
                if m in h:
                    f.write(f"| {m} | {r[h.index(m)]} | {units[h.index(m)]} |\n")
            f.write("\n")

def traffic(dst, *pairs):
    doc = {"source": "ncu --set full --clock-control none (one launch each, cold cache); see the matching profiles/*_full_*.md", "workloads": {}}
    num = lambda x: float(x.replace(",", "")) if x not in ("", "n/a") else None
    for pair in pairs:
        wl, src = pair.split("=", 1)
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h, units = rows[0], rows[1]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}

        def col(r, m):  # bytes in bytes, times in microseconds, everything else as printed
            if m not in h or num(r[h.index(m)]) is None:
                return None
            return num(r[h.index(m)]) * scale.get(units[h.index(m)], 1)
        w = doc["workloads"].setdefault(wl, {})
        for r in rows[2:]:
            name = re.sub(r"^void ", "", r[h.index("Kernel Name")].split("(")[0]).replace("rv::", "")
            name = re.sub(r"<\(bool\)([01])>", r"<\1>", name)
            if name in w:
                continue
            w[name] = {"dram_bytes_per_launch": (col(r, "dram__bytes_read.sum") or 0) + (col(r, "dram__bytes_write.sum") or 0),
                       "us_per_launch_under_ncu": col(r, "gpu__time_duration.sum"),
                       "grid": r[h.index("Grid Size")], "block": r[h.index("Block Size")],
                       "alu_pipe_pct": col(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                       "lsu_pipe_pct": col(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                       "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                       "sm_throughput_pct": col(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                       "dram_throughput_pct": col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                       "l2_bytes": col(r, "lts__t_bytes.sum")}
    with open(dst, "w") as f:
        json.dump(doc, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], *sys.argv[3:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
