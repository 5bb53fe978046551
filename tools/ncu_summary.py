"""Summarise an .ncu-rep (captured under gpurun) into a small markdown file under profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.md "title / command" """
import csv, subprocess, sys, io

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"]
lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on, one launch)", ""]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    lines += [f"## {name[:110]}", "", "| metric | value | unit |", "|---|---|---|"]
    for h, u, v in zip(hdr, units, r):
        if h in want:
            lines.append(f"| {h} | {v} | {u} |")
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    h2 = srows[1]; ci = {h: i for i, h in enumerate(h2)}
    data = []
    for r in srows[2:]:
        try: s = int(r[ci["# Samples"]])
        except Exception: continue
        st = {h: int(r[ci[h]]) for h in h2 if h.startswith("stall_") and "Not Issued" not in h and r[ci[h]].isdigit() and int(r[ci[h]]) > 0}
        data.append((s, r[ci["Source"]].strip(), st))
    tot = sum(d[0] for d in data) or 1
    lines += ["## hottest SASS instructions (warp-stall samples)", "", "| samples | % | instruction | top stall reasons |", "|---|---|---|---|"]
    for s, ins, st in sorted(data, key=lambda d: -d[0])[:14]:
        top = ", ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        lines.append(f"| {s} | {100*s/tot:.1f} | `{ins[:80]}` | {top} |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
