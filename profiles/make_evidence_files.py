"""Turns what profiles/collect_evidence.sh brought back (gpurun_out/ev/) into the tracked evidence files:
profiles/r2_launch_list.txt, profiles/r2d_dega_ncu_summary.txt, profiles/r2_traffic.json.  Runs here (needs ncu for reading the report)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = os.path.join(ROOT, "gpurun_out", "ev")
PR = os.path.join(ROOT, "profiles")


def launch_list():
    rows = [r for r in csv.reader(open(os.path.join(EV, "launches.csv"))) if len(r) >= 15 and r[0].isdigit()]
    L = [(r[4].split("(")[0].replace("void ", "").replace("dc::", ""), r[8], float(r[14].replace(",", "")) / 1e6) for r in rows]
    first_dec = next(i for i, x in enumerate(L) if x[0].startswith("validate_index"))
    a0 = first_dec - 6          # the step's encode in front of the first decode
    steps = L[a0:a0 + 63]       # 3 warm-up + 4 timed steps, 9 kernels each
    d = json.load(open(os.path.join(EV, "plain.json")))
    km = d["kernel_ms"]
    ks = ("front", "bac_encode", "scan", "gather", "bac_decode", "back")
    ksum = sum(km[k] for k in ks)
    out = []
    out.append('round 2, launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(front_|back_|bac_|scan_|gather_|publish_|validate_)" -c 600 --csv \\')
    out.append("    python bench.py --steps 4 --warmup 3 --no-cpu --e2e-stream-steps 4`   (profiles/collect_evidence.sh; the -k filter keeps the library's kernels:")
    out.append("    without it the first 600 launches are torch's generator of the synthetic input).  FINAL library of the round.")
    out.append("Cold-cache, serialised: compare SHARES, not absolutes.  The same command had exited 0 without ncu just before: %.2f ms/step, %.1f GB/s;" % (d["ms_per_step"], d["value"]))
    out.append("bench.py's own CUDA-event times of that run: " + ", ".join("%s %.3f" % (k, km[k]) for k in ks) + " ms")
    out.append("  = " + " / ".join("%.1f" % (100 * km[k] / ksum) for k in ks) + " % of the kernel time of a step (front / bac_encode / scan / gather / bac_decode / back)")
    out.append("")
    out.append("A. the device-resident steps (launches %d..%d of the list: 3 warm-up + 4 timed steps, 9 kernels each)" % (a0, a0 + 62))
    out.append("%-26s %9s %10s %8s" % ("kernel (dc::)", "launches", "mean ms", "share"))
    per = collections.OrderedDict()
    for n, g, ms in steps:
        per.setdefault(n, []).append(ms)
    tot = sum(sum(v) for v in per.values())
    for n, v in sorted(per.items(), key=lambda x: -sum(x[1])):
        out.append("%-26s %9d %10.3f %7.1f%%" % (n, len(v), sum(v) / len(v), 100 * sum(v) / tot))
    out.append("%-26s %9s %10.3f" % ("sum per step", "", tot / 7))
    out.append("")
    out.append("B. the rest of the 600 launches: the figures at other block lengths and the host-buffer (end-to-end) steps, whose passes run as half-wave chunks")
    out.append("%-26s %-14s %9s %10s" % ("kernel (dc::)", "grid", "launches", "mean ms"))
    per2 = collections.OrderedDict()
    for n, g, ms in L[:a0] + L[a0 + 63:]:
        per2.setdefault((n, g), []).append(ms)
    for (n, g), v in sorted(per2.items(), key=lambda x: -sum(x[1])):
        if sum(v) / len(v) >= 0.01:
            out.append("%-26s %-14s %9d %10.3f" % (n, g, len(v), sum(v) / len(v)))
    out.append("(grid 128 / 32: whole channels, 4096 streams; grid 148: the full passes at 1024-sample blocks and the half-wave chunks of the host-buffer steps, "
               "averaged together; the scan / publish / index-check kernels take 3-6 us each)")
    open(os.path.join(PR, "r2_launch_list.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:22]))


def ncu_files():
    rep = os.path.join(EV, "dega.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/ev_raw.csv", "w").write(raw)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    open("/tmp/ev_source.csv", "w").write(src)
    summ = subprocess.run([sys.executable, os.path.join(PR, "summarize.py"), "/tmp/ev_raw.csv"], capture_output=True, text=True).stdout
    mix = subprocess.run([sys.executable, os.path.join(PR, "summarize_source.py"), "/tmp/ev_source.csv", "140600000", "warp-symbol"], capture_output=True, text=True).stdout
    hdr = ("round 2, capture d (FINAL library of the round: encoder groups with the interval in one register, decoder with the lean dense-payload reader, back end with\n"
           "the medium path and its CTAs per SM picked per launch, gather with four loads in flight):\n"
           '`ncu --set full --clock-control none --import-source on -k regex:"^(bac_encode|bac_decode|front_dega|back_dega|gather)" -c 5 python profiles/prof_step.py 4096 1168 1`\n'
           "(profiles/collect_evidence.sh; the same command had exited 0 without ncu first).  workload: 4096 channels x 86400 samples, block_len 1168 -> 303104 streams;\n"
           "4.50 G coded symbols per pass = 140.6 M warp-symbols; 11.06 M warp-samples (32 samples each) per pass.\n\n")
    note = ("instruction mix from the source page (ncu -i ... --page source --csv --print-source sass), per warp-symbol (coder) -- for the front end, the gather and the\n"
            "back end multiply by 12.7 to read the figures per warp-sample:\n")
    open(os.path.join(PR, "r2d_dega_ncu_summary.txt"), "w").write(hdr + summ + note + mix)
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum per launch from the `ncu --set full` capture of the final library "
                       "(profiles/r2d_dega_ncu_summary.txt), workload 4096 x 86400, block_len 1168",
           "workload": {"channels": 4096, "samples": 86400, "block_len": 1168}, "warp_instructions": {}}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        d = dict(zip(h, r))
        u = dict(zip(h, units))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
        out[name] = int(sum(float(d[k].replace(",", "")) * mult[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
        out["warp_instructions"][name] = int(float(d["smsp__inst_executed.sum"].replace(",", "")))
        print(name, d["gpu__time_duration.sum"], d["smsp__issue_active.avg.pct_of_peak_sustained_active"], d["smsp__thread_inst_executed_per_inst_executed.ratio"])
    out["warp_instructions"]["source"] = "profiles/r2d_dega_ncu_summary.txt smsp__inst_executed.sum (the data is deterministic: same count every launch)"
    json.dump(out, open(os.path.join(PR, "r2_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    launch_list()
    ncu_files()
