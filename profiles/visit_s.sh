#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_r1t.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu_r1t.log
python profiles/k2_time.py 2>&1 | tail -1
python bench.py > $out/bench_r1t.json 2> $out/bench_r1t.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r1t.json"))
for k in ("value", "ms_per_step", "e2e", "e2e_resident", "lockstep_4096_placements_per_s", "lockstep_4096_cuda_graph_placements_per_s", "random_policy_placements_per_s_per_gpu", "e2e_lockstep_host_policy"):
    v = d.get(k); print(k, v if not isinstance(v, dict) else v.get("value"))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms_per_launch")})
PY
bash profiles/prof_kernel.sh k2_r1t k_step 1
