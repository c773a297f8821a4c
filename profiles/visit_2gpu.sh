#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $out/bench_2gpu_r1u.json 2> $out/bench_2gpu_r1u.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_2gpu_r1u.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 | cut -c1-200
