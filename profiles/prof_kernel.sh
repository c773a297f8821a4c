#!/bin/bash
# full ncu capture of one kernel launch of the prof_run.py workload:
#   bash profiles/prof_kernel.sh <tag> <kernel regex> [skip launches] [n_env] [T]
tag=$1; pat=$2; skip=${3:-1}; n=${4:-1048576}; T=${5:-8}
out=gpurun_out; mkdir -p $out
python profiles/prof_run.py $n $T > $out/plain_$tag.log 2>&1 || { echo "prof_run failed"; tail -5 $out/plain_$tag.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f -o $out/prof_$tag python profiles/prof_run.py $n $T > $out/ncu_$tag.log 2>&1
tail -3 $out/ncu_$tag.log
