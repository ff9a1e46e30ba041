#!/usr/bin/env python
"""Compact summary of an `ncu -i X.ncu-rep --page raw --csv` export: one row per metric, one column per profiled launch,
restricted to the counters profiles/README.md argues from (duration, DRAM bytes, DRAM / L2 / L1 / SM throughput, tensor pipe,
shared-memory wavefronts by source, issue slots, occupancy, registers, stall ratios).

  ncu -i gpurun_out/foo.ncu-rep --page raw --csv > /tmp/foo.csv
  python tools/ncu_summary.py /tmp/foo.csv profiles/ncu_foo_r02.csv "header comment ..."
"""
import csv
import re
import sys

KEEP = [
    r"^Kernel Name$", r"^Grid Size$", r"^Block Size$", r"^launch__cluster_size$", r"^launch__registers_per_thread$",
    r"^launch__occupancy_limit_", r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$",
    r"^dram__bytes_(read|write)\.sum\.per_second$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
    r"^sm__inst_executed_pipe_(tensor_subpipe_hmma|uniform|fma|alu|xu|lsu)\.avg\.pct_of_peak_sustained_active$",
    r"^l1tex__data_pipe_(tc|lsu)_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?$",
    r"^l1tex__data_pipe_lsu_wavefronts_mem_shared_op_(ld|st)\.sum$", r"^l1tex__t_sector_hit_rate\.pct$",
    r"^lts__t_sector_hit_rate\.pct$", r"^sm__cycles_active\.avg$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
    r"^smsp__average_warps?_(latency_)?issue_stalled_.*_per_issue_active\.ratio$",
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    comment = sys.argv[3:] if len(sys.argv) > 3 else []
    rows = list(csv.reader(open(src)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    keep = [i for i, h in enumerate(hdr) if any(re.search(p, h) for p in KEEP)]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        for c in comment:
            w.writerow(["# " + c])
        w.writerow(["metric", "unit"] + [f"launch {k}" for k in range(len(launches))])
        for i in keep:
            vals = [r[i] if i < len(r) else "" for r in launches]
            if hdr[i] == "Kernel Name":
                vals = [re.sub(r"\(.*", "", v).replace("void ", "").replace("dfcsa::<unnamed>::", "").replace("unnamed>::", "")[:60] for v in vals]
            if "issue_stalled" in hdr[i] and all((v in ("", "n/a") or abs(float(v)) < 0.05) for v in vals):
                continue
            w.writerow([hdr[i], units[i]] + vals)


if __name__ == "__main__":
    main()
