#!/bin/bash
# A/B of experimental builds on K2 / random rollout: VARIANTS="a b" bash profiles/ab_k2.sh  (gpurun_exp_<name>.so)
for v in ${VARIANTS:-0}; do echo "== variant $v"; TB_SO_PATH=/root/repo/gpurun_exp_$v.so python profiles/k2_time.py 2>&1 | tail -1; done
