#!/usr/bin/env python
"""Copy one gpu_round.sh visit's evidence from gpurun_out/ (scratch) into profiles/ (tracked) and refresh
k1_latest.json (the ncu figures bench.py quotes for `roofline.traffic` and `roofline.integer_pipe`).

    python profiles/collect_round.py <tag>
"""
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def main():
    tag = sys.argv[1]
    for k in ("k1", "k2", "k3", "k3r"):
        for t in ("lines", "stalls", "summary"):
            shutil.copy(os.path.join(OUT, "%s_%s_%s.txt" % (tag, k, t)), os.path.join(PROF, "%s_%s_%s.txt" % (tag, k, t)))
    for src, dst in (("launches_bench_%s.csv", "%s_launches_bench.csv"), ("launches_%s.csv", "%s_launches.csv"),
                     ("bench_%s.json", "%s_bench.json"), ("bench_ref_%s.json", "%s_bench_reference_arm.json"),
                     ("pytest_gpu_%s.log", "%s_pytest_gpu.log")):
        shutil.copy(os.path.join(OUT, src % tag), os.path.join(PROF, dst % tag))
    t = open(os.path.join(OUT, "%s_k1_summary.txt" % tag)).read()

    def g(name):
        return float(re.search(r"^\s*%s\s+([0-9.]+)" % re.escape(name), t, re.M).group(1))
    kernel = re.search(r":: void (k_afterstates<[^>]*>)", t).group(1).replace(" ", "")
    d = {"capture": tag, "kernel": kernel,
         "workload": "profiles/prof_run.py: 2^20 envs, 10x20, 7-piece, boards after greedy play (as in bench.py's roofline leg)",
         "dram_bytes_read": g("dram__bytes_read.sum") * 1e6, "dram_bytes_write": g("dram__bytes_write.sum") * 1e6,
         "gpu_time_us": g("gpu__time_duration.sum"),
         "alu_pipe_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
         "fma_pipe_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
         "lsu_pipe_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
         "xu_pipe_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
         "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "warp_instructions": g("smsp__inst_executed.sum"),
         "threads_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio")}
    json.dump(d, open(os.path.join(PROF, "k1_latest.json"), "w"), indent=1)
    print(json.dumps(d))


if __name__ == "__main__":
    main()
