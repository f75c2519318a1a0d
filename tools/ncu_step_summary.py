"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py:
per-kernel share of ONE steady-state training step (the launches between two
tgn::msg_build_kernel launches; the cycle is one full step, rotated)."""
import collections, csv, re, sys
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -3
detail = len(sys.argv) > 3
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
data = [(row[ki], float(row[vi].replace(',', '')), row[gi]) for row in r]
starts = [i for i, d in enumerate(data) if 'msg_build_kernel' in d[0]]
step = data[starts[which]:starts[which + 1]]
tot = sum(v for _, v, _ in step)
print(f"launches in step: {len(step)}   sum of kernel durations: {tot/1000:.1f} us (ncu: serialised, cold caches)")
agg = collections.OrderedDict()
for k, v, _ in step:
    k = re.sub(r'\(.*', '', k).replace('void ', '')[:64]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/1000:9.1f} us {100*v/tot:5.1f}%  x{n:3d}  {k}")
if detail:
    for i, (k, v, g) in enumerate(step):
        k = re.sub(r'\(.*', '', k).replace('void ', '')[:60]
        print(f"{i:3d} {v/1000:7.1f} us grid {g:>14s} {k}")
