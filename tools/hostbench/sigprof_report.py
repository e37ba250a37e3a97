"""sigprof.samples (written by a bench run with PROF=1) -> hottest source lines; for library / kernel frames the caller is shown."""
import collections, subprocess, sys
exe = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = [l.split() for l in open("sigprof.samples")]

def lines(addrs):
    out = subprocess.run(["addr2line", "-e", exe] + addrs, capture_output=True, text=True).stdout.splitlines()
    return [o.split(" ")[0].split("/")[-1] for o in out]

rip, c1 = lines([r[0] for r in rows]), lines([r[2] for r in rows])
h, tot = collections.Counter(), 0
for r, a, b in zip(rows, rip, c1):
    lib = a.startswith(("stl_", "new_allocator", "vector.tcc", "alloc_traits", "??"))
    h[(a, b if lib else "")] += int(r[1])
    tot += int(r[1])
print(f"{tot} samples")
for (a, b), v in h.most_common(top):
    print(f"{v:6d} {100 * v / tot:5.1f}%  {a}" + (f"   <- {b}" if b else ""))
