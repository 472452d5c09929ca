"""Turns `ncu --page raw --csv` output into the short per-kernel summaries kept under profiles/.

usage: python profiles/summarize.py raw.csv [launches.csv] > profiles/<round>_summary.txt
"""
import csv
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"Kernel Name{'':61s}{d['Kernel Name']}")
        for k in KEEP:
            if k in d:
                print(f"{k:72s}{d[k]:>16s} {unit.get(k, '')}")
        st = [(float(d[k]), k) for k in hdr if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and d[k]]
        print("  top stall reasons (warps per issue-active cycle):")
        for v, k in sorted(st, reverse=True)[:6]:
            name = k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
            print(f"    {name:30s} {v:.3f}")
        print()
    if len(sys.argv) > 2:
        per = {}
        for r in csv.DictReader(l for l in open(sys.argv[2]) if not l.startswith("==")):
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            v = v / 1e6 if r.get("Metric Unit") in ("ns", "nsecond") else v / 1e3 if r.get("Metric Unit") in ("us", "usecond") else v
            per.setdefault(r["Kernel Name"].split("(")[0], []).append(v)
        tot = sum(sum(v) for v in per.values())
        print(f"{'kernel':40s} {'launches':>8s} {'mean ms':>10s} {'share':>8s}")
        for k, v in per.items():
            print(f"{k:40s} {len(v):8d} {sum(v) / len(v):10.3f} {100 * sum(v) / tot:7.1f}%")


if __name__ == "__main__":
    main()
