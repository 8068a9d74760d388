"""Summarise `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` : hottest CUDA lines.

usage: ncu_src.py dump.csv [top] [sort: stall|inst|smem]
"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
key = sys.argv[3] if len(sys.argv) > 3 else "stall"
rows = list(csv.reader(open(path)))
hdr = None
fname = ""
agg = []
for r in rows:
    if r and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ci = {}
        for i, h in enumerate(hdr):
            ci.setdefault(h, i)
        continue
    if hdr is None or len(r) < 20 or not r[0].strip().isdigit():
        continue

    def f(name):
        try:
            return float(r[ci[name]].replace(",", "") or 0)
        except (ValueError, KeyError, IndexError):
            return 0.0
    agg.append((f("Instructions Executed"), f("Warp Stall Sampling (All Samples)"), f("L1 Wavefronts Shared"),
                f("L1 Wavefronts Shared Excessive"), f("Avg. Threads Executed"), fname, r[0], r[1].strip()[:95]))
tot = sum(a[0] for a in agg) or 1
tots = sum(a[1] for a in agg) or 1
totw = sum(a[2] for a in agg) or 1
print(f"total warp-inst {tot:.3e}  stall samples {tots:.0f}  smem wavefronts {totw:.3e}")
print(" inst%  stall%  smemwf% (excess%) thr | file:line | source")
k = {"stall": 1, "inst": 0, "smem": 2}[key]
for n, st, wf, ex, thr, fn, ln, src in sorted(agg, key=lambda a: -a[k])[:top]:
    print(f"{n / tot * 100:6.2f} {st / tots * 100:7.2f} {wf / totw * 100:7.2f} ({ex / totw * 100:5.2f}) {thr:4.1f} | {fn}:{ln} | {src}")
