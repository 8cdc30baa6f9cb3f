#!/usr/bin/env python3
"""Top source lines of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py REPORT KERNEL_REGEX [N [LAUNCH_SKIP]]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
skip = sys.argv[4] if len(sys.argv) > 4 else "0"      # n-th launch matching the name (template instances share a base name)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern,
                      "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, data = "", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
    elif hdr and len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            ie = int(r[hdr["Instructions Executed"]]); smp = int(r[hdr["# Samples"]] or 0)
            te = int(r[hdr["Thread Instructions Executed"]])
        except ValueError:
            continue
        st = {k: int(r[i] or 0) for k, i in hdr.items() if k.startswith("stall_") and "Not Issued" not in k and r[i] not in ("", "-")}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        data.append((ie, smp, te, cur_file, r[0], r[1].strip()[:100], top))
ti = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("total warp-inst %d  samples %d  avg threads/inst %.1f" % (ti, ts, sum(d[2] for d in data) / max(ti, 1)))
for d in sorted(data, reverse=True)[:topn]:
    print("%5.1f%% inst %5.1f%% smp thr %4.1f | %s:%s | %s | %s" % (100 * d[0] / ti, 100 * d[1] / max(ts, 1), d[2] / max(d[0], 1), d[3], d[4], d[5],
                                                          " ".join("%s=%d" % (k[6:], v) for k, v in d[6])))
