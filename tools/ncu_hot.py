"""Development aid: hottest SASS instructions of one kernel from an .ncu-rep (first instance)."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.015
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hidx = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"]
h = rows[hidx[0]]
body = rows[hidx[0] + 1:(hidx[1] - 1 if len(hidx) > 1 else len(rows))]
body = [r for r in body if len(r) == len(h)]
si, ie, so = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
tot = sum(int(r[si]) for r in body)
print("kernel", rows[hidx[0] - 1][1][:80] if hidx[0] else "", "samples", tot, "instrs", len(body), "warp-instr executed", sum(int(r[ie]) for r in body))
names = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
for k, r in enumerate(body):
    if int(r[si]) > tot * thr:
        st = {n.replace("stall_", ""): r[h.index(n)] for n in names if r[h.index(n)] != "0"}
        print(k, r[so].strip()[:62], "| smp", r[si], "exec", r[ie], st)
