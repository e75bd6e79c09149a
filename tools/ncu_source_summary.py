"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` by source line:
warp-stall samples, executed instructions, shared-memory wavefronts.  Usage:
    python tools/ncu_source_summary.py file.csv [topN]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
cur_file, hdr = None, None
agg = defaultdict(lambda: defaultdict(float))
src_text = {}
stall_cols = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_lg", "stall_dispatch"]
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    key = (cur_file, int(r[0]))
    src_text[key] = r[1].strip()[:110]
    a = agg[key]

    def f(name):
        try:
            return float(d.get(name, "0") or 0)
        except ValueError:
            return 0.0

    a["samples"] += f("# Samples")
    a["inst"] += f("Instructions Executed")
    a["wf"] += f("L1 Wavefronts Shared")
    a["wf_ideal"] += f("L1 Wavefronts Shared Ideal")
    for c in stall_cols:
        a[c] += f(c)
tot = defaultdict(float)
for a in agg.values():
    for k, v in a.items():
        tot[k] += v
print(f"total samples {tot['samples']:.0f}  instructions {tot['inst']:.3e}  smem wavefronts {tot['wf']:.3e} (ideal {tot['wf_ideal']:.3e})")
print("stall mix: " + "  ".join(f"{c[6:]}={100*tot[c]/max(tot['samples'],1):.1f}%" for c in stall_cols))
byfile = defaultdict(float)
for (fl, ln), a in agg.items():
    byfile[fl] += a["samples"]
print("by file:", {k: f"{100*v/tot['samples']:.1f}%" for k, v in byfile.items()})
print(f"{'file:line':28s} {'smp%':>6s} {'inst%':>6s} {'wf%':>6s}  top stalls | source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((a[c], c[6:]) for c in stall_cols), reverse=True)[:2]
    sts = ",".join(f"{n}:{100*v/max(a['samples'],1):.0f}" for v, n in st)
    print(f"{key[0][:20]+':'+str(key[1]):28s} {100*a['samples']/tot['samples']:6.2f} {100*a['inst']/tot['inst']:6.2f} {100*a['wf']/max(tot['wf'],1):6.2f}  {sts:24s} | {src_text[key]}")
