#!/bin/bash
# BASELINE configs[4]: board-size sweep (10x20, 10x10, 6x12) on N GPUs of one box, greedy rollouts with the NCCL
# episode-statistics reduction, per-shape kernel rooflines from rank 0.  Under gpurun --gpus N:
#   bash profiles/board_sweep.sh <tag> <N> [boards...]
tag=$1; n=$2; shift 2; boards=${@:-10x20 10x10 6x12}
out=gpurun_out/board_sweep_${tag}_${n}gpu.json; : > $out
for b in $boards; do
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --board $b --steps 10 --warmup 3 --extras light >> $out 2>> gpurun_out/board_sweep_$tag.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --board $b --steps 10 --warmup 3 --extras light >> $out 2>> gpurun_out/board_sweep_$tag.err; fi
done
python - $out <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    r, k3 = d.get("roofline", {}), d.get("roofline_step_kernel", {})
    print("%-6s N=%d  %.3e placements/s  %.2f ms/step | K1 %.3f ms, HBM frac %.3f | K3 %.3e afterstates/s" % (
        d["config"]["board"], d["n_gpus"], d["value"], d["ms_per_step"], r.get("ms_per_launch", 0), r.get("frac", 0),
        k3.get("afterstates_scored_per_s") or 0))
PY
