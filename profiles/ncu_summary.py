#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics the roofline discussion uses, for every kernel in a capture.

    python profiles/ncu_summary.py <report.ncu-rep> [more reports ...]
"""
import csv
import re
import subprocess
import sys

WANT = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|launch__registers_per_thread|"
                  r"launch__shared_mem_per_block_static|launch__occupancy_limit_(registers|shared_mem|warps)|"
                  r"launch__grid_size|launch__block_size|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
                  r"smsp__thread_inst_executed_per_inst_executed\.ratio|"
                  r"sm__inst_executed_pipe_(alu|fma|xu|lsu|uniform|fmaheavy|fmalite)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|"
                  r"smsp__average_warp_latency_issue_stalled_.*|sm__cycles_active\.avg)$")


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print("== %s :: %s" % (rep, d["Kernel Name"][:90]))
            for k, u in zip(hdr, units):
                if WANT.match(k):
                    v = d[k]
                    if "stalled" in k:
                        try:
                            if float(v) < 0.05:
                                continue
                        except ValueError:
                            pass
                    print("   %-92s %s %s" % (k, v, u))


if __name__ == "__main__":
    main()
