"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel into markdown.
usage: python tools/launch_list.py gpurun_out/launches.csv profiles/rNN_launches_bench_step.md "command line" """
import csv, re, sys
from collections import OrderedDict

src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ci = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    val = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
    name = re.sub(r"\(.*$", "", r[ci["Kernel Name"]]).replace("void ", "").replace("gb::", "").strip()
    rows.append((name, us))


def table(rs):
    agg = OrderedDict()
    for n, us in rs:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(v[1] for v in agg.values())
    t = ["| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        t.append(f"| {n} | {c} | {us:.0f} | {us / c:.1f} | {100 * us / tot:.1f}% |")
    return t, tot


t, tot = table(rows)
md = ["# ncu launch list of ONE bench step, aggregated per kernel", "", f"command: `{cmd}`",
      "(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)", "",
      f"launches: {len(rows)}, sum of kernel time: {tot / 1000:.1f} ms", ""] + t
# one decode step = the launches between two consecutive decode_advance kernels
# (the last kernel of a step: reduce_head_argmax_cluster_kernel since the one-launch tail, decode_advance_kernel before it)
idx = [i for i, (n, _) in enumerate(rows) if n.startswith("reduce_head_argmax_cluster")] or \
      [i for i, (n, _) in enumerate(rows) if n.startswith("decode_advance")]
if len(idx) >= 3:
    seg = rows[idx[-2] + 1: idx[-1] + 1]
    t2, tot2 = table(seg)
    md += ["", f"## one decode step (graph nodes): {len(seg)} launches, {tot2 / 1000:.2f} ms serialised", ""] + t2
open(out, "w").write("\n".join(md) + "\n")
print(f"{len(rows)} launches -> {out}")
