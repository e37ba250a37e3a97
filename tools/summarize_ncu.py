"""Turns gpurun_out ncu artefacts into the small tracked summaries under profiles/.
usage: summarize_ncu.py launches <launches.csv> <out.md>   |   summarize_ncu.py full <file.ncu-rep> <out.md>"""
import collections, csv, subprocess, sys

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

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
