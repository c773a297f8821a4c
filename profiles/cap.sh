#!/bin/bash
# Full ncu captures of the prof_run.py workload, summarised on the box:  bash profiles/cap.sh <tag> k1 k3 k2 k3r ...
tag=$1; shift; out=gpurun_out; mkdir -p $out
python profiles/prof_run.py > $out/plain_$tag.log 2>&1 || { echo "prof_run failed"; tail -5 $out/plain_$tag.log; exit 1; }
cap() {  # name, kernel regex, skip, mangled-name substring for the line tools
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $out/prof_$1_$tag python profiles/prof_run.py > $out/ncu_$1_$tag.log 2>&1
  python profiles/ncu_summary.py $out/prof_$1_$tag.ncu-rep > $out/${tag}_$1_summary.txt 2>&1
  python profiles/ncu_lines.py $out/prof_$1_$tag.ncu-rep $4 60 > $out/${tag}_$1_lines.txt 2>&1
  python profiles/ncu_stalls.py $out/prof_$1_$tag.ncu-rep $4 > $out/${tag}_$1_stalls.txt 2>&1
  [ "$KEEP_REP" = "$1" ] || rm -f $out/prof_$1_$tag.ncu-rep
}
for k in "$@"; do
  case $k in
    k1) cap k1 k_afterstates 1 k_afterstatesILi10ELi20ELi0ELi256ELi3ELi256E ;;
    k3) cap k3 k_rollout_greedy 2 k_rollout_greedyILi10ELi20ELi256ELi3ELi256E ;;
    k3g) cap k3g k_rollout_greedy 5 k_rollout_greedyILi10ELi20ELi256ELi3ELi256E ;;   # 8 steps in the greedy steady state
    k2) cap k2 k_step 1 k_stepILi10ELi20ELi256ELi4E ;;
    k3r) cap k3r k_rollout_random 1 k_rollout_randomILi10ELi20EE ;;
  esac
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_$tag.csv python profiles/prof_run.py > $out/ncu_l_$tag.log 2>&1
echo done
