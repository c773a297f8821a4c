#!/bin/bash
# One GPU visit: parity tests, the bench line, the ncu launch list and full captures of K1 / K3.
# Usage (from the repo root, under gpurun):  bash profiles/gpu_round.sh [tag]
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cat $out/bench_$tag.json | cut -c1-600
# launch list of the bench command itself (shares of the step, not absolutes)
python bench.py --steps 2 --warmup 3 --no-extras > $out/bench_plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 3 --no-extras > $out/ncu_lb_$tag.log 2>&1
python profiles/prof_run.py > $out/plain_$tag.log 2>&1 || { echo "prof_run failed"; tail -5 $out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_$tag.csv python profiles/prof_run.py > $out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_afterstates -s 1 -c 1 -f -o $out/prof_k1_$tag python profiles/prof_run.py > $out/ncu_k1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_greedy -s 1 -c 1 -f -o $out/prof_k3_$tag python profiles/prof_run.py > $out/ncu_k3_$tag.log 2>&1
echo done
