#!/bin/bash
# A/B of experimental builds (gpurun_exp_<name>.so at the repo root, built with -DTB_ONLY_10x20 -DTB_EXP_<name>)
for v in ${VARIANTS:-0}; do echo "== variant $v"; TB_SO_PATH=/root/repo/gpurun_exp_$v.so K1CFGS="0" K3CFGS="0" bash profiles/quick.sh 2>&1 | grep -E "K1 cfg|K3 cfg"; done
