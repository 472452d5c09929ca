"""Instruction mix, active-lane buckets and stall-sample shares per kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`.

usage: python profiles/summarize_source.py source.csv UNITS [UNIT_NAME]   (UNITS = work items per launch, e.g. warp-symbols)"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    unit_name = sys.argv[3] if len(sys.argv) > 3 else "unit"
    kern, data, hdr = None, {}, None
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            kern, hdr = r[1], None
            data[kern] = []
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if kern and hdr and len(r) == len(hdr):
            data[kern].append(dict(zip(hdr, r)))
    for k, v in data.items():
        tot_inst = sum(int(d["Instructions Executed"]) for d in v)
        tot_samp = sum(int(d["# Samples"]) for d in v) or 1
        per = f" = {tot_inst / units:.1f} per {unit_name}" if units else ""
        print(f"{k}: warp instructions {tot_inst}{per}; stall samples {tot_samp}")
        lanes = collections.defaultdict(float)
        by = collections.defaultdict(lambda: [0, 0])
        for d in v:
            n = int(d["Instructions Executed"])
            t = float(d["Avg. Threads Executed"]) if n else 0
            lanes["1-2 lanes" if t <= 2 else "3-8" if t <= 8 else "9-24" if t <= 24 else "25-32"] += n
            parts = d["Source"].split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "SHF")) else op.split(".")[0]
            by[op][0] += n
            by[op][1] += int(d["# Samples"])
        print("  warp instructions by active lanes: " + ", ".join(f"{a}: {100 * b / tot_inst:.1f} %" for a, b in sorted(lanes.items())))
        for op, (n, s) in sorted(by.items(), key=lambda x: -x[1][0])[:18]:
            per = f"{n / units:7.2f} per {unit_name}  " if units else ""
            print(f"   {op:12s} {per}{100 * n / tot_inst:6.2f} % of instructions  {100 * s / tot_samp:6.2f} % of stall samples")
        print()


if __name__ == "__main__":
    main()
