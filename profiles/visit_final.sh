#!/bin/bash
# final visit of the round: parity suite, bench line, launch lists, full captures of K1 / K3 / K2
tag=${1:-r1v}
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cut -c1-400 $out/bench_$tag.json
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref_$tag.json 2>/dev/null; cut -c1-200 $out/bench_ref_$tag.json
python bench.py --steps 2 --warmup 3 --no-extras > $out/bench_plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 3 --no-extras > $out/ncu_lb_$tag.log 2>&1
python profiles/prof_run.py > $out/plain_$tag.log 2>&1 || { echo "prof_run failed"; tail -5 $out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_$tag.csv python profiles/prof_run.py > $out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_afterstates -s 1 -c 1 -f -o $out/prof_k1_$tag python profiles/prof_run.py > $out/ncu_k1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_greedy -s 2 -c 1 -f -o $out/prof_k3_$tag python profiles/prof_run.py > $out/ncu_k3_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 1 -c 1 -f -o $out/prof_k2_$tag python profiles/prof_run.py > $out/ncu_k2_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_random -s 1 -c 1 -f -o $out/prof_k3r_$tag python profiles/prof_run.py > $out/ncu_k3r_$tag.log 2>&1
echo done
