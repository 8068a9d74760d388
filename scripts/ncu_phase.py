"""Instruction share per source-line range of one kernel.
usage: ncu_phase.py src.csv file.cuh 'name:marker' 'name:marker' ...   (ranges run from marker to next marker)"""
import csv, collections, sys
path, fname = sys.argv[1], sys.argv[2]
marks = [m.split(':', 1) for m in sys.argv[3:]]
agg = collections.Counter(); thr = collections.Counter(); stall = collections.Counter()
cur = ''; hdr = None; kern = 0
for r in csv.reader(open(path)):
    if r and r[0] == 'Kernel Name':
        kern += 1
    if kern > 1:
        break
    if r and r[0] in ('File Path', 'File Name'):
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = {h: i for i, h in reversed(list(enumerate(r)))}; continue
    if hdr is None or len(r) < 20 or not r[0].strip().isdigit():
        continue
    def f(n):
        try: return float(r[hdr[n]] or 0)
        except Exception: return 0.0
    k = (cur, int(r[0]))
    agg[k] += f('Instructions Executed'); thr[k] += f('Thread Instructions Executed'); stall[k] += f('Warp Stall Sampling (All Samples)')
tot = sum(agg.values()); ts = sum(stall.values()) or 1
lines = open([p for p in [fname] ][0]).read().split('\n')
base = fname.split('/')[-1]
pos = []
for name, mk in marks:
    ln = next(i + 1 for i, l in enumerate(lines) if mk in l and not l.lstrip().startswith('//   '))
    pos.append((name, ln))
pos.append(('END', 10**9))
print(f"total warp-inst {tot:.4g}")
for (name, a), (_, b) in zip(pos, pos[1:]):
    s = sum(v for (f_, l), v in agg.items() if f_ == base and a <= l < b)
    t = sum(v for (f_, l), v in thr.items() if f_ == base and a <= l < b)
    st = sum(v for (f_, l), v in stall.items() if f_ == base and a <= l < b)
    print(f"{name:14s} lines {a:4d}-  inst {s / tot * 100:6.2f}%  stall {st / ts * 100:6.2f}%  thr/inst {t / max(s, 1):5.1f}")
o = sum(v for (f_, l), v in agg.items() if f_ != base)
print(f"other files    inst {o / tot * 100:6.2f}%")
