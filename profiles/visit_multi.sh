#!/bin/bash
# One multi-GPU visit (under gpurun --gpus N):  bash profiles/visit_multi.sh <tag> <N>
# NCCL parity test, the bench line at N GPUs (whole-job e2e included), BASELINE configs[4] board sweep at N GPUs.
tag=$1; n=$2; out=gpurun_out; mkdir -p $out
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -2
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n > $out/bench_${tag}_${n}gpu.json 2> $out/bench_${tag}_${n}gpu.err; echo "bench rc=$?"
grep -m3 -E "NVLS|comm .* nranks|Init COMPLETE" $out/bench_${tag}_${n}gpu.err | cut -c1-200
python - $out/bench_${tag}_${n}gpu.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step")}, "e2e", d["e2e"]["value"], d.get("episode_stats_invariants"))
PY
bash profiles/board_sweep.sh $tag $n
