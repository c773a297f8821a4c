#!/bin/bash
# r1q visit: full GPU parity suite, the bench line (new e2e), small-batch A/B
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_r1q.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu_r1q.log
python bench.py > $out/bench_r1q.json 2> $out/bench_r1q.err; echo "bench rc=$?"; tail -3 $out/bench_r1q.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r1q.json"))
for k in ("value", "ms_per_step", "e2e", "e2e_unpipelined", "e2e_resident", "lockstep_4096_placements_per_s", "lockstep_4096_cuda_graph_placements_per_s"):
    print(k, d.get(k))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms_per_launch")})
PY
python profiles/small_batch.py > $out/small_batch_r1q.json 2>&1; cat $out/small_batch_r1q.json
