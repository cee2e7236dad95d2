"""Extracts what the hardware executed for one launch from a .ncu-rep and stores it under profiles/executed.json, where
bench.py's `roofline.executed` reads it: thread-level FP32 and FP64 FLOPs (FMA = 2) and warp instructions of the render kernel.
usage: python profiles/ncu_executed.py <report.ncu-rep> <workload> [launches_per_frame]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, workload = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[2]
    d = dict(zip(hdr, vals))

    def g(k):
        return float(d[k].replace(",", ""))

    # thread-level (predicated-on) instruction counts per SASS opcode, from the report's source page
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    shdr = srows[1]
    ia, it = shdr.index("Source"), shdr.index("Predicated-On Thread Instructions Executed")
    thr = {}
    for r in srows[2:]:
        t = r[ia].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        thr[op] = thr.get(op, 0) + int(r[it])
    fp32 = 2 * thr.get("FFMA", 0) + thr.get("FMUL", 0) + thr.get("FADD", 0)
    fp64 = 2 * thr.get("DFMA", 0) + thr.get("DMUL", 0) + thr.get("DADD", 0)
    entry = {"kernel": d.get("Kernel Name"), "fp32_flop": fp32, "fp64_flop": fp64, "warp_inst": g("smsp__inst_executed.sum"),
             "gpu_time_us_under_ncu": g("gpu__time_duration.sum") * (1e3 if rows[1][hdr.index("gpu__time_duration.sum")] == "ms" else 1.0),
             "dram_bytes": (g("dram__bytes_read.sum"), rows[1][hdr.index("dram__bytes_read.sum")], g("dram__bytes_write.sum"),
                            rows[1][hdr.index("dram__bytes_write.sum")]),
             "source": os.path.relpath(rep, ROOT)}
    path = os.path.join(ROOT, "profiles", "executed.json")
    allx = json.load(open(path)) if os.path.exists(path) else {}
    allx[workload] = entry
    json.dump(allx, open(path, "w"), indent=1, sort_keys=True)
    print(workload, entry)


if __name__ == "__main__":
    main()
