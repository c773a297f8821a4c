#!/bin/bash
out=gpurun_out; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}"; }
run 29521 > $out/bench_8gpu_r1w.json 2> $out/bench_8gpu_r1w.err; echo "bench rc=$?"
run 29522 --board 10x10 --steps 10 --no-extras > $out/sweep8_10x10_r1w.json 2>/dev/null; echo "rc=$?"
run 29523 --board 6x12 --steps 10 --no-extras > $out/sweep8_6x12_r1w.json 2>/dev/null; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_8gpu_r1w", "sweep8_10x10_r1w", "sweep8_6x12_r1w"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["n_gpus"], d["ms_per_step"], (d.get("e2e") or {}).get("value"))
    except Exception as ex:
        print(f, "failed", ex)
PY
