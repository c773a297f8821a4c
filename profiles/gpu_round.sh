#!/bin/bash
# One GPU visit: parity tests, the bench line + reference arm, the ncu launch lists, and full captures of K1 / K3 / K2 /
# the random rollout summarised ON the box (the reports together exceed gpurun's 64 MiB return limit; KEEP_REP=k1 keeps
# one).  Usage (from the repo root, under gpurun):  bash profiles/gpu_round.sh [tag]
# Afterwards, here:  python profiles/collect_round.py <tag>   (copies the evidence into profiles/, refreshes *_latest.json)
tag=${1:-r2}
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; cut -c1-300 $out/bench_$tag.json
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref_$tag.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-extras > $out/bench_plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 3 --no-extras > $out/ncu_lb_$tag.log 2>&1
bash profiles/cap.sh $tag k1 k3 k3g k2 k3r
ls -la $out | head -50
echo done
