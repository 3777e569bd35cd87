"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of scripts/step_once.py: the last training step
and the inference forward, grouped by kernel.  usage: launch_summary.py launches.csv [top_n]"""
import collections, csv, re, sys

rows = []
for line in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')):
    if line[0] == "ID":
        continue
    name = re.sub(r"^void ", "", line[4])
    name = re.sub(r"\(.*", "", name).replace("lpm::", "")
    rows.append((name, line[8], float(line[-1]) / 1000.0))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
starts = [i for i, r in enumerate(rows) if r[0].startswith(("sample_stats_kernel", "sample_stats_warp_kernel", "step_begin_kernel"))]
opt = [i for i, r in enumerate(rows) if r[0].startswith(("rank_adam_kernel", "rank_adam_tile_kernel", "mt_adam_kernel"))]
b = opt[-1] + 1
a = max(i for i in starts if i < opt[-1] and rows[i][0].startswith(("step_begin_kernel", "sample_stats")) and
        (rows[i][0].startswith("step_begin_kernel") or not any(rows[j][0].startswith("step_begin_kernel") for j in range(max(0, i - 3), i))))
while b < len(rows) and rows[b][0].startswith(("transpose_2d_kernel", "split_hi_lo_kernel", "rank_adam", "mt_")) or \
        (b < len(rows) and "elementwise_kernel" in rows[b][0] and b + 1 < len(rows) and rows[b + 1][0].startswith("split_hi_lo_kernel")):
    b += 1      # refresh of the transposed centre shadows and of the split-precision weight operands


def table(title, seg):
    agg = collections.OrderedDict()
    for n, g, t in seg:
        k = (n, g)
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"== {title}: {len(seg)} launches, {tot:.1f} us")
    for (n, g), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {t:8.1f} us {100 * t / tot:5.1f}%  x{c:3d}  {n[:70]} {g}")
    fam = collections.defaultdict(float)
    for n, g, t in seg:
        fam[re.sub(r"<.*", "", n)] += t
    print("families: " + ", ".join(f"{k} {v:.0f}" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:14]))


table("train step", rows[a:b])
# the LAST inference forward (steady state): it starts with input_bn's finalize right before the last sample_apply
last_apply = max(i for i, r in enumerate(rows) if r[0].startswith("sample_apply"))
table("inference forward", rows[last_apply - 1:])
