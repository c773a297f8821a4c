#!/usr/bin/env python
"""bench.py -- placements/sec of the batched Tetris environment (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the reference on host cores

Workload (config.workload): BASELINE.json configs[3]'s per-GPU shard -- 2^20 envs per GPU, 10x20 board, 7-piece
set, greedy linear policy with the BCTS weights of game.py:111-118, game-over detection + auto-reset, synthetic
(seeded, in-kernel RNG) piece streams.  A "step" is one fused rollout launch of `rollout_steps` placements for every
env (= example_play.py:11-21 loop bodies) followed, for N > 1, by the NCCL all-reduce of the episode statistics.
Weak scaling: envs per GPU fixed, env ids global (rank r owns [r*E, (r+1)*E)), no per-step communication.
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C, R, PIECE_SET = 10, 20, 1     # --board overrides C x R (BASELINE configs[4]: 10x20, 10x10, 6x12)
STATE_READ_BYTES = 64          # 3 x 128-bit row planes + 128-bit meta per env (include/tetris_b200.h)
METRIC = "placements/sec (env steps/sec), greedy-linear policy, whole job"
# The reference itself is pure Python and does not travel to the GPU box (no /root/reference there), so its own speed
# rides along as a labelled constant: measured in the build container (SURVEY.md section 6 / BASELINE.md section 2).
PYTHON_REFERENCE = {
    "value": 200.0, "unit": "placements/s", "cores": 1, "kind": "reference (Python, unmodified), constant -- not timed in this run",
    "provenance": "SURVEY.md section 6: corrected example_play.py loop, 10x20, 7-piece set, greedy BCTS argmax, 300 "
                  "placements, 1 core of the build container's 8-vCPU Xeon, NumPy 2.3.5 / CPython 3.12.3 "
                  "(random policy: 170 placements/s; afterstate feature vectors: 3.2-4.7 k/s)",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--rollout-steps", type=int, default=32, help="placements per env per step")
    ap.add_argument("--seed", type=int, default=0x5EED)
    ap.add_argument("--board", default="10x20", help="columns x rows: 10x20 (headline), 10x10, 6x12, 8x16")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="stream-pipelined chunks of the host-buffer e2e cycle")
    ap.add_argument("--no-extras", action="store_true", help="skip roofline / cpu baseline / e2e side measurements")
    ap.add_argument("--extras", default="full", choices=["full", "light", "none"],
                    help="light: only the per-kernel rooflines beside the timed step (board sweeps)")
    return ap.parse_args()


def config_of(args, n_gpus):
    return {
        "workload": "BASELINE configs[3] per-GPU shard: greedy linear policy (BCTS weights, game.py:111-118) playing "
                    "with game-over + auto-reset, %dx%d board, 7-piece set; %d envs/GPU x %d placements per step"
                    % (C, R, args.envs, args.rollout_steps),
        "board": "%dx%d" % (C, R), "piece_set": "7-piece (game.py:41-47)", "policy": "greedy-linear BCTS",
        "envs_per_gpu": args.envs, "total_envs": args.envs * n_gpus, "rollout_steps": args.rollout_steps,
        "l2": "flushed between timed steps (256 MiB write outside the timed events)",
        "parallelism": "envs sharded over %d GPU(s); stats all-reduce only" % n_gpus,
    }


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (every 20 ms).  nvidia-smi block-buffers its
    output into the pipe, so lines arrive in bursts: the samples of the timed region are selected by nvidia-smi's own
    timestamp against the wall clock at mark_begin / mark_end, not by when they arrived."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc, self.thread = index, [], None, None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        self.t1 = datetime.datetime.now()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    @staticmethod
    def parse(line):
        """(time, sm MHz, max sm MHz, power W, [reason flags]) of one csv line, or None."""
        f = [x.strip() for x in line.split(",")]
        if len(f) < 10:
            return None
        try:
            ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
            return ts, float(f[2]), float(f[3]), float(f[4]), [x.lower().startswith("active") for x in f[6:10]]
        except ValueError:
            return None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)                                     # let the sample of the last 20 ms be written
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=5)                      # drain what was still buffered in the pipe
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [r for r in (self.parse(ln) for ln in list(self.lines)) if r is not None]
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        where = "timed region"
        if not inside and self.t0 is not None:               # a region shorter than the sampling period: nearest samples
            inside = sorted(rows, key=lambda r: abs((r[0] - self.t0).total_seconds()))[:3]
            where = "nearest to the timed region"
        sm = sorted(r[1] for r in inside)
        reasons = sorted({names[k] for r in inside for k in range(4) if r[4][k]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in inside), default=None),
                "power_w_max": max((r[3] for r in inside), default=None), "samples": len(sm), "sampled": where,
                "reasons": reasons}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def issue_ruler(kernel_key, sm_count, sm_hz):
    """The integer roof of the tile kernels, as an executed-work ruler: they are bit manipulation, so what bounds them is the
    rate at which the 4 schedulers of every SM can issue warp-instructions -- roof = SMs x 4 x clock.  Executed
    instructions cannot be counted without a profiler, so instruction count AND duration both come from ONE ncu
    capture (profiles/<kernel>_latest.json, written by profiles/collect_round.py), never a stored count over a live
    time; the capture records the hash of the kernel sources and the block says when the sources have changed since.
      issue_frac  = warp-instructions / duration / roof              (how busy the issue slots are with whatever issued)
      useful_frac = issue_frac x (active threads per instruction / 32)   (lanes that did work)"""
    if (C, R) != (10, 20):
        return None                              # the captures are of the headline board
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "%s_latest.json" % kernel_key)))
    except Exception:
        return None
    from tetris_b200 import _lib
    roof = sm_count * 4 * sm_hz
    rate = prof["warp_instructions"] / (prof["gpu_time_us"] * 1e-6)
    out = {
        "roof_warp_instructions_per_s": roof, "roof_source": "%d SMs x 4 schedulers x %.0f MHz" % (sm_count, sm_hz / 1e6),
        "capture": prof.get("capture"), "capture_kernel": prof.get("kernel"), "capture_workload": prof.get("workload"),
        "capture_stale": prof.get("source_hash") != _lib.source_hash(),
        "warp_instructions_per_launch": prof["warp_instructions"], "capture_ms": prof["gpu_time_us"] * 1e-3,
        "issue_frac": rate / roof, "active_threads_per_instruction": prof["threads_per_instruction"],
        "useful_frac": rate / roof * prof["threads_per_instruction"] / 32.0,
        "issue_active_pct_ncu": prof.get("issue_active_pct"), "alu_pipe_pct_ncu": prof.get("alu_pipe_pct"),
        "xu_pipe_pct_ncu": prof.get("xu_pipe_pct"), "lsu_pipe_pct_ncu": prof.get("lsu_pipe_pct"),
        "warps_active_pct_ncu": prof.get("warps_active_pct"), "registers_per_thread": prof.get("registers_per_thread"),
        "shared_bank_conflict_frac": (prof["shared_bank_conflict_wavefronts"] / prof["shared_wavefronts"])
        if prof.get("shared_wavefronts") else None,
        "dram_bytes_per_launch_ncu": (prof.get("dram_bytes_read") or 0) + (prof.get("dram_bytes_write") or 0),
    }
    return out


# ------------------------------------------------------------------------------------------------------
def cpu_port_rate(seconds_target, threads, policy=1, seed=1):
    """The oracle port of the reference's loop on `threads` host cores; returns (placements/s, afterstates/s, sample)."""
    from oracle import oracle as orc
    n_env = 256 * threads
    b = orc.Batch(C, R, n_env, piece_set=PIECE_SET, seed=seed)
    b.reset()
    b.rollout(30, 0, threads=threads)                       # realistic boards, untimed
    t0 = time.perf_counter()
    st = b.rollout(4, policy, threads=threads)
    probe = time.perf_counter() - t0
    rate = n_env * 4 / max(probe, 1e-6)
    T = max(4, int(seconds_target * rate / n_env))
    t0 = time.perf_counter()
    st = b.rollout(T, policy, threads=threads)
    dt = time.perf_counter() - t0
    sample = "%d envs x %d placements (greedy BCTS, %dx%d, 7-piece) in %.1f s on %d threads" % (n_env, T, C, R, dt, threads)
    return n_env * T / dt, float(st[4]) / dt, sample


def run_reference(args):
    """CPU arm.  The reference is pure Python and cannot travel to the GPU box (/root/reference is absent there), so
    this times the oracle's C port of its loop (oracle/tetris_oracle.c, pinned to the reference by golden fixtures)
    on every host core.  Each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    n_env = 512 * threads
    T = args.rollout_steps
    b = orc.Batch(C, R, n_env, piece_set=PIECE_SET, seed=args.seed)
    b.reset()
    b.rollout(30, 0, threads=threads)
    for _ in range(args.warmup):
        b.rollout(T, 1, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b.rollout(T, 1, threads=threads)
    dt = time.perf_counter() - t0
    value = n_env * T * args.steps / dt
    sample = "%d envs x %d placements per step on %d host threads (C port of the reference loop)" % (n_env, T, threads)
    cfg = config_of(args, args.gpus)
    # same workload (board, pieces, policy, placements per step); the env count is this arm's own bounded sample -- a
    # CPU's rate does not depend on it -- and is stated as such instead of the GPU arm's 2^20 per GPU
    cfg.update({"envs_per_gpu": None, "total_envs": n_env, "envs_cpu_sample": n_env,
                "workload": cfg["workload"].replace("%d envs/GPU" % args.envs, "%d envs (bounded CPU sample)" % n_env),
                "l2": "n/a (CPU)", "parallelism": "%d host threads, envs split evenly" % threads})
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "placements/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "placements/s", "cores": threads, "kind": "port", "sample": sample,
                         "python_reference": PYTHON_REFERENCE},
        "e2e": {"value": value, "unit": "placements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from tetris_b200 import BCTS_WEIGHTS, BatchedTetris, _lib
    from tetris_b200.distributed import reduce_stats

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU product path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG is the caller's (the driver reads the communicator lines); its output goes to stderr unless the
        # caller says otherwise, because stdout carries the one JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    K, W, T, E = args.steps, args.warmup, args.rollout_steps, args.envs
    W = max(W, 3)

    env = BatchedTetris(C, R, E, piece_set=PIECE_SET, seed=args.seed, env_offset=rank * E, device=dev)
    env.rollout(30, "random")                                # realistic, de-synchronised boards (untimed)
    env.stats.zero_()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    weights = np.asarray(BCTS_WEIGHTS, np.float32)

    reduced = torch.zeros_like(env.stats)

    def step():
        env.rollout(T, "greedy", weights)
        if world > 1:
            return reduce_stats(env.stats, out=reduced)      # end-of-rollout reduction of episode statistics: ONE
        return env.stats                                     # NCCL all-gather + a one-warp combine kernel

    sampler = ClockSampler(local)
    sampler.start()                                          # nvidia-smi takes a moment to start: begin before warm-up
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()                                     # only samples taken during the timed region count
    wall0 = time.perf_counter()
    for s, e in ev:
        flush.zero_()
        s.record()
        step()
        e.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    sampler.mark_end()
    clocks = sampler.stop()
    ms = sum(s.elapsed_time(e) for s, e in ev)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    placements = float(world) * E * T * K
    value = placements / (ms * 1e-3)
    # The reduced episode statistics of the whole job (all ranks, warm-up + timed steps; the statistics were zeroed
    # after the untimed random warm-up), with the integer invariants every correct run satisfies -- evidence that all
    # ranks took part in the reduction and that every env made every placement.
    final = (reduced if world > 1 else env.stats).cpu().tolist()
    red = dict(zip(_lib.STATS, final))
    want = world * E * T * (W + K)
    if red["placements"] != want or sum(red["lines%d" % i] for i in range(5)) != red["placements"] \
            or red["lines"] != sum(i * red["lines%d" % i] for i in range(5)) \
            or red["reward"] != red["lines"] - red["placements"] - 100 * red["episodes"]:
        raise SystemExit("bench.py: reduced episode statistics violate their invariants: %r (expected %d placements)"
                         % (red, want))

    out = {
        "metric": METRIC, "value": value, "unit": "placements/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": config_of(args, world),
        "clocks": clocks, "gpu_launches": K * (2 if world > 1 else 1), "wall_s_timed_region": wall,
        "episode_stats_reduced": {k: red[k] for k in _lib.STATS if not k.startswith("reserved")},
        "episode_stats_invariants": "ok: placements == world x envs x rollout_steps x (warmup + steps) = %d; "
                                    "lines0..4 sum to placements; lines and reward follow from the histogram" % want,
        "dtype_note": "u32 column bit masks; features / scores leave the kernels as float32 (exact small integers, half-integers)",
    }

    # ---- e2e through the public API with HOST buffers (every rank; whole-job value = all ranks' envs / max time).
    # tetris_b200.HostRollout.play: every step the boards and pieces of all envs come from pinned host memory (H2D +
    # tb_import_boards), are played T placements by the fused rollout, and go back to the host (tb_export_boards + D2H)
    # together with the episode statistics -- the cycle of a caller that keeps its games on the host.  Copies, the two
    # conversion kernels and the final synchronisation are inside the timed region.  The env range is cut into chunks on
    # separate streams so the copies of one chunk overlap the rollout kernel of another.
    if args.no_extras:
        args.extras = "none"
    light = args.extras == "light"
    e2e = e2e_serial = None
    if args.extras == "full":
        from tetris_b200 import HostRollout

        def time_host_rollout(chunks):
            hr = HostRollout(C, R, E, chunks=chunks, piece_set=PIECE_SET, seed=args.seed, env_offset=rank * E, device=dev)
            for sub in hr.envs:
                sub.rollout(30, "random")
            hr.pull()
            for _ in range(2):
                hr.play(T, "greedy", weights)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                hr.play(T, "greedy", weights)                # synchronous: host buffers valid on return
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(world) * E * T * K / float(dt.item()), hr.h2d_bytes, hr.d2h_bytes

        v, h2d, d2h = time_host_rollout(args.e2e_chunks)
        e2e = {"value": v, "unit": "placements/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "chunks": args.e2e_chunks,
               "note": "HostRollout.play per step and rank: boards + pieces of all envs H2D from pinned memory -> "
                       "tb_import_boards -> tb_rollout (T placements per env) -> tb_export_boards -> boards, heights, "
                       "pieces and statistics D2H, synchronised; %d chunks on separate streams overlap copies with the "
                       "kernel; whole-job value (all ranks, max time over ranks); bytes are per rank" % args.e2e_chunks}
        if world == 1:
            v1, _, _ = time_host_rollout(1)
            e2e_serial = {"value": v1, "unit": "placements/s", "chunks": 1,
                          "note": "the same cycle as one chunk on one stream (no copy/compute overlap)"}

    if rank == 0 and args.extras != "none":
        stats = env.stats_dict()
        out["rollout_afterstates_per_s_per_gpu"] = None
        # afterstates evaluated inside the timed rollouts (this rank): measured, not estimated
        torch.cuda.synchronize()
        a0 = env.stats_dict()["afterstates"]
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); env.rollout(T, "greedy", weights); e.record(); torch.cuda.synchronize()
        out["rollout_afterstates_per_s_per_gpu"] = (env.stats_dict()["afterstates"] - a0) / (s.elapsed_time(e) * 1e-3)
        out["episode_stats_rank0"] = {k: stats[k] for k in ("placements", "episodes", "lines", "max_ep_lines")}
        props = torch.cuda.get_device_properties(dev)
        sm_hz = 1e6 * float((json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("sm_max_mhz", 1965.0))
                            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1965.0)
        peak, peak_src = measured_peak_gbs()
        # ---- the timed step's own kernel (K3, k_rollout_greedy).  Its HBM traffic is the env state in and out once per
        # launch of T placements, so the HBM fraction is tiny by construction; what bounds it is instruction issue (the
        # work is K1's enumeration + features fused with the policy and the env step): see issue_ruler.
        k3_bytes = 2 * E * (STATE_READ_BYTES + 8)
        # k3g: captured in the greedy steady state (the boards this timed region plays on); k3: boards after random play
        k3_issue = issue_ruler("k3g", props.multi_processor_count, sm_hz) or issue_ruler("k3", props.multi_processor_count, sm_hz)
        out["roofline_step_kernel"] = {
            "kernel": "k_rollout_greedy<%d,%d> (K3): %d placements per env per launch, %d envs" % (C, R, T, E),
            "bound": "issue slots (integer / bit work); HBM sees the env state once per launch",
            "ms_per_launch": ms / K, "placements_per_s": E * T / (ms / K * 1e-3),
            "afterstates_scored_per_s": out["rollout_afterstates_per_s_per_gpu"],
            "hbm": {"achieved": k3_bytes / (ms / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": k3_bytes / (ms / K * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": k3_bytes,
                    "traffic": (k3_issue or {}).get("dram_bytes_per_launch_ncu"), "peak_source": peak_src},
            "issue": k3_issue,
            "issue_on_random_play_boards": issue_ruler("k3", props.multi_processor_count, sm_hz),
        }

        # ---- roofline of the afterstate kernel (K1, the north_star's roofline target), timed live
        feats = torch.empty((E, env.a_max, 8), dtype=torch.float32, device=dev)
        valid = torch.empty(E, dtype=torch.int64, device=dev)
        count = torch.empty(E, dtype=torch.int32, device=dev)
        for _ in range(3):
            env.get_after_states(out=(feats, valid, count))
        torch.cuda.synchronize()
        k1 = []
        for _ in range(5):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); env.get_after_states(out=(feats, valid, count)); e.record()
            torch.cuda.synchronize()
            k1.append(s.elapsed_time(e))
        k1_ms = sum(k1) / len(k1)
        n_slots = torch.as_tensor([_lib.lib().tb_num_slots(p, C) for p in range(9)], device=dev)
        slots_total = int(n_slots[env.export_boards()[2].long()].sum().item())     # placements enumerated
        rows_written = int(count.sum().item())                                      # legal ones: a 32 B feature row each
        # algorithmic bytes: state read + mask/count written + 32 B per legal afterstate
        alg_bytes = E * (STATE_READ_BYTES + 12) + 32 * rows_written
        achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
        k1_issue = issue_ruler("k1", props.multi_processor_count, sm_hz)
        out["roofline"] = {
            "kernel": "k_afterstates<%d,%d> (K1: enumerate + 8 features for every placement of %d envs)" % (C, R, E),
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (k1_issue or {}).get("dram_bytes_per_launch_ncu"),
            "traffic_source": "ncu --set full capture %s (profiles/k1_latest.json), dram__bytes_read+write per launch on the "
                              "profiles/prof_run.py boards%s" % ((k1_issue or {}).get("capture"),
                                                                 "; STALE: kernel sources changed since" if (k1_issue or {}).get("capture_stale") else ""),
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
            "ms_per_launch": k1_ms, "afterstates_per_s": slots_total / (k1_ms * 1e-3),
            "legal_afterstates_per_s": rows_written / (k1_ms * 1e-3),
            "issue": k1_issue,
            "note": "HBM is the roof the contract asks for and the lower one on paper; the kernel is bound by instruction "
                    "issue (DESIGN.md section 4): `frac` is the HBM fraction, `issue` the executed-work ruler",
        }
        # the compact output format (int16 = 2 x feature): 16 B per legal afterstate
        try:
            feats16 = torch.empty((E, env.a_max, 8), dtype=torch.int16, device=dev)
            for _ in range(2):
                env.get_after_states(out=(feats16, valid, count), compact=True)
            k1c = []
            for _ in range(5):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); env.get_after_states(out=(feats16, valid, count), compact=True); e.record()
                torch.cuda.synchronize()
                k1c.append(s.elapsed_time(e))
            k1c_ms = sum(k1c) / len(k1c)
            algc = E * (STATE_READ_BYTES + 12) + 16 * rows_written
            out["roofline"]["compact_i16"] = {"ms_per_launch": k1c_ms, "algorithmic_bytes_per_launch": algc,
                                              "achieved": algc / (k1c_ms * 1e-3) / 1e9, "frac": algc / (k1c_ms * 1e-3) / 1e9 / peak}
            del feats16
        except Exception as ex:
            out["roofline"]["compact_i16"] = {"error": repr(ex)}
        # K2 (tb_step), timed live on the same boards
        try:
            a0 = torch.zeros(E, dtype=torch.int32, device=dev)
            saved = env.state.clone()
            env.step(a0, auto_reset=True, check=False)
            k2 = []
            for _ in range(5):
                env.state.copy_(saved)
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); env.step(a0, auto_reset=True, check=False); e.record()
                torch.cuda.synchronize()
                k2.append(s.elapsed_time(e))
            env.state.copy_(saved)
            k2_ms = sum(k2) / len(k2)
            k2_bytes = E * (2 * STATE_READ_BYTES + 4 + 4 + 1 + 4 + 32)
            out["roofline_k2"] = {"kernel": "k_step<%d,%d> (K2)" % (C, R), "bound": "hbm", "ms_per_launch": k2_ms,
                                  "algorithmic_bytes_per_launch": k2_bytes, "achieved": k2_bytes / (k2_ms * 1e-3) / 1e9,
                                  "peak": peak, "unit": "GB/s", "frac": k2_bytes / (k2_ms * 1e-3) / 1e9 / peak,
                                  "issue": issue_ruler("k2", props.multi_processor_count, sm_hz)}
            del saved
        except Exception as ex:
            out["roofline_k2"] = {"error": repr(ex)}
        del feats

        if not light:
            out["e2e"] = e2e
            if e2e_serial is not None:
                out["e2e_unpipelined"] = e2e_serial
            host_stats = torch.empty(len(_lib.STATS), dtype=torch.int64).pin_memory()
            # the same API with the games resident on the device (only weights in, statistics out)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(K):
                env.rollout(T, "greedy", np.array(BCTS_WEIGHTS, np.float32))
                host_stats.copy_(env.stats, non_blocking=True)
                torch.cuda.synchronize()
            out["e2e_resident"] = {"value": E * T * K / (time.perf_counter() - t0), "unit": "placements/s",
                                   "h2d_bytes_per_step": 32, "d2h_bytes_per_step": 8 * len(_lib.STATS),
                                   "note": "BatchedTetris.rollout() with device-resident games: weights in, statistics out"}

            # ---- e2e of the lockstep API a host-side policy uses: features to the host, actions back (PCIe-bound)
            try:
                nl = min(E, 1 << 18)
                env2 = BatchedTetris(C, R, nl, piece_set=PIECE_SET, seed=args.seed + 1, device=dev)
                env2.rollout(20, "random")
                hc = torch.empty(nl, dtype=torch.int32).pin_memory()
                ha = torch.zeros(nl, dtype=torch.int32).pin_memory()
                ho = torch.empty((nl, 8), dtype=torch.float32).pin_memory()
                n_it = 3
                for compact in (False, True):                              # float32 rows, then the int16 (2 x feature) rows
                    hf = torch.empty((nl, env2.a_max, 8), dtype=torch.int16 if compact else torch.float32).pin_memory()
                    df = torch.empty((nl, env2.a_max, 8), dtype=hf.dtype, device=dev)
                    dv = torch.empty(nl, dtype=torch.int64, device=dev)
                    dc = torch.empty(nl, dtype=torch.int32, device=dev)
                    for it in range(n_it + 1):                             # first iteration = warm-up (pinned pages, caches)
                        if it == 1:
                            torch.cuda.synchronize()
                            t0 = time.perf_counter()
                        f, v, c = env2.get_after_states(out=(df, dv, dc), compact=compact)
                        hf.copy_(f, non_blocking=True); hc.copy_(c, non_blocking=True)
                        torch.cuda.synchronize()
                        obs, rew, done, lines = env2.step(ha.to(dev, non_blocking=True), auto_reset=True, check=False)
                        ho.copy_(obs, non_blocking=True)
                        torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    out["e2e_lockstep_host_policy" + ("_compact" if compact else "")] = {
                        "value": nl * n_it / dt, "unit": "placements/s", "envs": nl,
                        "h2d_bytes_per_step": 4 * nl, "d2h_bytes_per_step": nl * (env2.a_max * (16 if compact else 32) + 4 + 32),
                        "note": "get_after_states -> features D2H -> actions H2D -> step -> obs D2H (action 0 for all envs)"
                                + ("; features as int16 = 2 x feature (TB_FLAG_FEATS_I16)" if compact else "")}
                    del hf, df
                del env2
            except Exception as ex:                                        # side measurement only
                out["e2e_lockstep_host_policy"] = {"error": repr(ex)}

            # ---- random policy and the 4096-env lockstep config, for context
            try:
                envr = BatchedTetris(C, R, E, piece_set=PIECE_SET, seed=args.seed + 2, device=dev)
                envr.rollout(30, "random")
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); envr.rollout(64, "random"); e.record(); torch.cuda.synchronize()
                out["random_policy_placements_per_s_per_gpu"] = E * 64 / (s.elapsed_time(e) * 1e-3)
                del envr
                env4 = BatchedTetris(C, R, 4096, piece_set=PIECE_SET, seed=args.seed, device=dev)
                g = torch.Generator(device=dev); g.manual_seed(0)
                for it in range(60):
                    if it == 10:
                        torch.cuda.synchronize(); t0 = time.perf_counter()
                    f, v, c = env4.get_after_states()
                    a = (torch.randint(0, 2 ** 31 - 1, (4096,), device=dev, generator=g) % c.long()).int()
                    env4.step(a, auto_reset=True, check=False)
                torch.cuda.synchronize()
                out["lockstep_4096_placements_per_s"] = 4096 * 50 / (time.perf_counter() - t0)
                # the same iteration captured once in a CUDA graph (launch-bound at this batch size)
                envg = BatchedTetris(C, R, 4096, piece_set=PIECE_SET, seed=args.seed, device=dev)
                replay, _outs = envg.capture_lockstep(           # default CUDA generator: graph-safe philox offsets
                    lambda f, v, c: (torch.randint(0, 2 ** 31 - 1, (4096,), device=dev) % c.long().clamp(min=1)).int())
                for _ in range(10):
                    replay()
                torch.cuda.synchronize(); t0 = time.perf_counter()
                for _ in range(200):
                    replay()
                torch.cuda.synchronize()
                out["lockstep_4096_cuda_graph_placements_per_s"] = 4096 * 200 / (time.perf_counter() - t0)
                # the same 4096 envs with the policy inside the kernel (fused rollouts, small-batch tile configuration)
                for pol in ("random", "greedy"):
                    envg.rollout(64, pol)
                    torch.cuda.synchronize()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(); envg.rollout(256, pol); e.record(); torch.cuda.synchronize()
                    out["fused_4096_%s_placements_per_s" % pol] = 4096 * 256 / (s.elapsed_time(e) * 1e-3)
            except Exception as ex:
                out["random_policy_placements_per_s_per_gpu"] = repr(ex)

            # ---- CPU baseline beside it (N = 1 only): the oracle port on this box's host cores
            if world == 1:
                threads = os.cpu_count() or 1
                v, av, sample = cpu_port_rate(12.0, threads)
                out["cpu_baseline"] = {"value": v, "unit": "placements/s", "cores": threads, "kind": "port",
                                       "sample": sample, "afterstates_per_s": av, "python_reference": PYTHON_REFERENCE}

    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    C, R = (int(x) for x in a.board.lower().split("x"))
    STATE_READ_BYTES = 16 * ((R + 4 + 7) // 8 + 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
