#!/bin/bash
# One GPU visit: parity tests, the bench line + reference arm, the ncu launch lists, and full captures of K1 / K3 / K2 /
# the random rollout summarised ON the box (the four reports together exceed gpurun's 64 MiB return limit; only K1's
# report travels back).  Usage (from the repo root, under gpurun):  bash profiles/gpu_round.sh [tag]
tag=${1:-r2}
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cut -c1-300 $out/bench_$tag.json
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref_$tag.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-extras > $out/bench_plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 3 --no-extras > $out/ncu_lb_$tag.log 2>&1
python profiles/prof_run.py > $out/plain_$tag.log 2>&1 || { echo "prof_run failed"; tail -5 $out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_$tag.csv python profiles/prof_run.py > $out/ncu_l_$tag.log 2>&1
cap() {  # name, kernel regex, skip, mangled-name substring for the line tools
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $out/prof_$1_$tag python profiles/prof_run.py > $out/ncu_$1_$tag.log 2>&1
  python profiles/ncu_summary.py $out/prof_$1_$tag.ncu-rep > $out/${tag}_$1_summary.txt 2>&1
  python profiles/ncu_lines.py $out/prof_$1_$tag.ncu-rep $4 > $out/${tag}_$1_lines.txt 2>&1
  python profiles/ncu_stalls.py $out/prof_$1_$tag.ncu-rep $4 > $out/${tag}_$1_stalls.txt 2>&1
}
cap k1 k_afterstates 1 k_afterstatesILi10ELi20ELb0ELi256ELi3ELi256E
cap k3 k_rollout_greedy 2 k_rollout_greedyILi10ELi20ELi256ELi2ELi256E
cap k2 k_step 1 k_stepILi10ELi20E
cap k3r k_rollout_random 1 k_rollout_randomILi10ELi20E
rm -f $out/prof_k3_$tag.ncu-rep $out/prof_k2_$tag.ncu-rep $out/prof_k3r_$tag.ncu-rep
ls -la $out | head -40
echo done
