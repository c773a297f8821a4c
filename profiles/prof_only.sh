#!/bin/bash
# ncu captures only: launch list + full captures of K1 / K3 (prof_run.py workload).  bash profiles/prof_only.sh <tag> [n_env] [T]
tag=${1:-x}; n=${2:-1048576}; T=${3:-8}
out=gpurun_out; mkdir -p $out
python profiles/prof_run.py $n $T > $out/plain_$tag.log 2>&1 || { echo "prof_run failed"; tail -5 $out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_$tag.csv python profiles/prof_run.py $n $T > $out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_afterstates -s 1 -c 1 -f -o $out/prof_k1_$tag python profiles/prof_run.py $n $T > $out/ncu_k1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_greedy -s 1 -c 1 -f -o $out/prof_k3_$tag python profiles/prof_run.py $n $T > $out/ncu_k3_$tag.log 2>&1
grep -h "tb::" $out/launches_$tag.csv | awk -F'","' '{print $5, $NF}' | cut -c1-160
