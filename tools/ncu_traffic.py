"""profiles/r02_dram_traffic.json from the ncu report of tools/decode_traffic.py: dram__bytes_read.sum + dram__bytes_write.sum per launch
for every decode-step kernel class, next to the algorithmic bytes (bench.py reads the per-class figures for `roofline.traffic`).
    python tools/ncu_traffic.py gpurun_out/r2_decode_traffic.ncu-rep profiles/r02_dram_traffic.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
units = rows[1]
def val(r, name):
    v = float(r[ci[name]].replace(",", "")); u = units[ci[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "nsecond": 1e-3, "ms": 1e3}.get(u, 1)
launches = []
for r in rows[2:]:
    launches.append({"kernel": r[ci["Kernel Name"]][:60], "grid": r[ci["Grid Size"]] if "Grid Size" in ci else None,
                     "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"), "us": val(r, "gpu__time_duration.sum")})
names = ["qkv", "o", "gu", "down", "head", "attention"]
Hd, I, V, B, H, ctx = 4096, 11008, 32114, 16, 32, 1030
alg = {"qkv": 3 * Hd * Hd * 2, "o": Hd * Hd * 2, "gu": 2 * I * Hd * 2, "down": Hd * I * 2, "head": V * Hd * 2, "attention": 2 * B * H * (ctx + 1) * 128 * 2}
per = {}
for n, l in zip(names, launches):
    per[n] = dict(l, traffic=l["dram_read"] + l["dram_write"], algorithmic=alg[n], ratio=(l["dram_read"] + l["dram_write"]) / alg[n])
counts = {"qkv": 32, "o": 32, "gu": 32, "down": 32, "head": 1}
gemm = sum(per[n]["traffic"] * c for n, c in counts.items()) / 129.0
res = {"source": rep, "per_launch": per, "gemm_bf16_tcgen05_kernel<16>": gemm, "decode_attention_tma_kernel": per["attention"]["traffic"],
       "note": "traffic = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (ncu --set full, cold L2 after a 512 MB flush); the GEMM figure is the "
               "launch-count-weighted mean of the five shapes of a step (32 x qkv, o, gate/up, down + 1 x heads) / 129, comparable with roofline.algorithmic_bytes_per_launch; "
               "writes are the fp32 split-K partials (implementation traffic)"}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: (round(v["ratio"], 3), round(v["us"], 1)) for k, v in per.items()}))
