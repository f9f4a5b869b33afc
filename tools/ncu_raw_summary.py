"""Summarise an `ncu --page raw --csv` export (one block per kernel launch): the metrics DESIGN.md / profiles/README.md quote.
Usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_raw_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name", "?"))
    for w in want:
        if w in d:
            print("   %-72s %s %s" % (w, d[w], units[hdr.index(w)]))
    top = sorted(((float(d[h]), h) for h in stall if d[h] not in ("", "n/a")), reverse=True)[:6]
    print("   stall cycles per issued instruction: " + ", ".join(
        "%s %.2f" % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v) for v, h in top))
    print()
