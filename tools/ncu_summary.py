#!/usr/bin/env python3
"""Summarise an .ncu-rep (one `ncu --set full` capture) into JSON: per kernel duration, DRAM traffic, cache hit rates,
issue utilisation, occupancy.  usage: ncu_summary.py REPORT OUT.json [TRAFFIC.json [BUILD_ID]]
TRAFFIC.json (kernel short name -> dram bytes read+written per launch, + build_id of the library that was profiled) is what
bench.py reports as roofline.traffic -- only when the library it runs has that build id."""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
WANT = {
    "duration": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
    "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct", "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "warp_inst": "smsp__inst_executed.sum", "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3}
res, traffic = [], {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
    d = {"kernel": name}
    for k, m in WANT.items():
        if m in ix and r[ix[m]] not in ("", "n/a"):
            v = float(r[ix[m]].replace(",", ""))
            u = units[ix[m]]
            d[k] = v * SCALE.get(u, 1)
    if "duration" in d:
        d["duration_ms"] = d.pop("duration")
    res.append(d)
    if "dram_read" in d:
        short = name.replace("_kernel", "").replace("void ", "")
        short = {"bpe_lookup": "bpe_encode", "pretok_split16": "pretok_split", "bpe_encode_pieces<1>": "long_scan", "bpe_encode_pieces<2>": "bpe_encode_fused"}.get(short, short)   # bench.py's names
        traffic[short] = d["dram_read"] + d.get("dram_write", 0)
json.dump(res, open(out, "w"), indent=1)
if len(sys.argv) > 3:
    if len(sys.argv) > 4:
        traffic["build_id"] = sys.argv[4]
    json.dump(traffic, open(sys.argv[3], "w"), indent=1)
for d in res:
    print(d)
