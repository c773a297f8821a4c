#!/usr/bin/env python
"""Copy one GPU visit's evidence from gpurun_out/ (scratch) into profiles/ (tracked) and refresh <kernel>_latest.json
(the ncu figures bench.py quotes next to its live timings: DRAM traffic, issue-slot and pipe utilisation, executed
instructions, shared-memory wavefronts).  Every <kernel>_latest.json records the hash of the kernel sources it was
captured from (tetris_b200._lib.source_hash); bench.py flags the block as stale when the sources have changed since.

    python profiles/collect_round.py <tag> [k1 k3 k2 k3r]
"""
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
WORKLOAD = {
    "k1": "profiles/prof_run.py: tb_afterstates over 2^20 envs, 10x20, 7-piece, boards after greedy play (bench.py's roofline leg)",
    "k3": "profiles/prof_run.py: tb_rollout greedy, 8 placements per env over 2^20 envs, 10x20, 7-piece, boards after random play",
    "k3g": "profiles/prof_run.py: tb_rollout greedy, 8 placements per env over 2^20 envs, 10x20, 7-piece, boards in the greedy steady state (what bench.py's timed region sees)",
    "k2": "profiles/prof_run.py: tb_step (action 0) over 2^20 envs, 10x20, 7-piece, boards after greedy play",
    "k3r": "profiles/prof_run.py: tb_rollout random, 8 placements per env over 2^20 envs, 10x20, 7-piece",
}


def parse_summary(path):
    t = open(path).read()

    def g(name, to=None):
        m = re.search(r"^\s*%s\s+([0-9.]+)\s*(\S*)" % re.escape(name), t, re.M)
        if not m:
            return None
        v = float(m.group(1))
        if to is not None:
            v *= UNIT.get(m.group(2), 1.0)
        return v
    kernel = re.search(r":: void ((?:tb::)?k_\w+<[^>]*>)", t).group(1).replace(" ", "")
    pct = "avg.pct_of_peak_sustained_active"
    return {
        "kernel": kernel,
        "dram_bytes_read": g("dram__bytes_read.sum", "byte"), "dram_bytes_write": g("dram__bytes_write.sum", "byte"),
        "gpu_time_us": g("gpu__time_duration.sum", "us"),
        "registers_per_thread": g("launch__registers_per_thread"), "grid_size": g("launch__grid_size"),
        "warps_active_pct": g("sm__warps_active." + pct),
        "alu_pipe_pct": g("sm__inst_executed_pipe_alu." + pct), "fma_pipe_pct": g("sm__inst_executed_pipe_fma." + pct),
        "lsu_pipe_pct": g("sm__inst_executed_pipe_lsu." + pct), "xu_pipe_pct": g("sm__inst_executed_pipe_xu." + pct),
        "issue_active_pct": g("smsp__issue_active." + pct),
        "warp_instructions": g("smsp__inst_executed.sum"),
        "threads_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "shared_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "shared_bank_conflict_wavefronts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    }


def main():
    from tetris_b200 import _lib
    tag = sys.argv[1]
    kernels = sys.argv[2:] or ["k1", "k3", "k3g", "k2", "k3r"]
    for k in kernels:
        for t in ("lines", "stalls", "summary"):
            src = os.path.join(OUT, "%s_%s_%s.txt" % (tag, k, t))
            if os.path.exists(src):
                shutil.copy(src, os.path.join(PROF, "%s_%s_%s.txt" % (tag, k, t)))
        summ = os.path.join(OUT, "%s_%s_summary.txt" % (tag, k))
        if os.path.exists(summ):
            d = {"capture": tag, "source_hash": _lib.source_hash(), "workload": WORKLOAD.get(k, "")}
            d.update(parse_summary(summ))
            json.dump(d, open(os.path.join(PROF, "%s_latest.json" % k), "w"), indent=1)
            print(k, json.dumps(d))
    for src, dst in (("launches_bench_%s.csv", "%s_launches_bench.csv"), ("launches_%s.csv", "%s_launches.csv"),
                     ("bench_%s.json", "%s_bench.json"), ("bench_ref_%s.json", "%s_bench_reference_arm.json"),
                     ("pytest_gpu_%s.log", "%s_pytest_gpu.log"), ("ab_%s.txt", "%s_ab.txt")):
        if os.path.exists(os.path.join(OUT, src % tag)):
            shutil.copy(os.path.join(OUT, src % tag), os.path.join(PROF, dst % tag))


if __name__ == "__main__":
    main()
