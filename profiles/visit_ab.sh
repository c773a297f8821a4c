#!/bin/bash
# A/B visit: baseline library (profiles/ab_builds/libtb_<base>.so) against the current build on the same harness,
# then full captures of K1 / K3 of the current build.   bash profiles/visit_ab.sh <tag> <base>
tag=$1; base=$2; out=gpurun_out; mkdir -p $out
for rep in 1 2; do
  TB_SO_PATH=$PWD/profiles/ab_builds/libtb_$base.so python profiles/ab_cfg.py --k1 0 --k3 0 > $out/ab_${tag}_base$rep.json 2>> $out/ab_$tag.err
  python profiles/ab_cfg.py --k1 0,6 --k3 0,6 > $out/ab_${tag}_new$rep.json 2>> $out/ab_$tag.err
done
cat $out/ab_${tag}_base*.json $out/ab_${tag}_new*.json
cap() {  # name, kernel regex, skip, mangled-name substring for the line tools
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $out/prof_$1_$tag python profiles/prof_run.py > $out/ncu_$1_$tag.log 2>&1
  python profiles/ncu_summary.py $out/prof_$1_$tag.ncu-rep > $out/${tag}_$1_summary.txt 2>&1
  python profiles/ncu_lines.py $out/prof_$1_$tag.ncu-rep $4 60 > $out/${tag}_$1_lines.txt 2>&1
  python profiles/ncu_stalls.py $out/prof_$1_$tag.ncu-rep $4 > $out/${tag}_$1_stalls.txt 2>&1
}
if [ "${3:-cap}" = "cap" ]; then
cap k1 k_afterstates 1 k_afterstatesILi10ELi20ELb0ELi256ELi3ELi256E
cap k3 k_rollout_greedy 2 k_rollout_greedyILi10ELi20ELi256ELi2ELi256E
rm -f $out/prof_k3_$tag.ncu-rep
fi
echo done
