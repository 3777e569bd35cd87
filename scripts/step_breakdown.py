"""Aggregate an ncu launch list (gpu__time_duration.sum csv) of scripts/step_once.py into a per-kernel table of the
last training step and the inference forward."""
import csv, re, sys, collections
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
names = [re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('lpm::', '')[:60] for r in rows]
idx = [i for i, n in enumerate(names) if 'sample_stats' in n]
s = idx[-1]
e = max(i for i, n in enumerate(names) if 'adam' in n.lower())
def table(lo, hi, title):
    agg = collections.OrderedDict()
    tot = 0.0
    for n, r in zip(names[lo:hi], rows[lo:hi]):
        us = float(r['Metric Value']) / 1e3
        tot += us
        a = agg.setdefault((n, r['Grid Size']), [0, 0.0]); a[0] += 1; a[1] += us
    print(f"== {title}: {hi - lo} launches, {tot:.1f} us")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:8.1f} us {100 * v[1] / tot:5.1f}%  x{v[0]:3d}  {k[0]} {k[1]}")
    fam = {}
    for n, r in zip(names[lo:hi], rows[lo:hi]):
        f = n.split('<')[0]
        fam[f] = fam.get(f, 0) + float(r['Metric Value']) / 1e3
    print("families:", ", ".join(f"{k} {v:.0f}" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:14]))
table(s, e + 1, "train step")
table(e + 1, len(rows), "inference forward")
