#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_r1r.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu_r1r.log
for cfg in 4 5; do TB_K1_CFG=$cfg TB_K3_CFG=$cfg python -m pytest tests/test_cuda_parity.py -x -q -k "lockstep or rollout_vs or ragged or afterstates_fixture or trace or fuzz" > $out/pytest_cfg${cfg}_r1r.log 2>&1; echo "pytest cfg$cfg rc=$?"; tail -2 $out/pytest_cfg${cfg}_r1r.log; done
python profiles/k2_time.py 2>&1 | tail -1
python profiles/small_batch.py > $out/small_batch_r1r.json 2>&1; cat $out/small_batch_r1r.json
python bench.py --steps 5 --no-extras 2>/dev/null | cut -c1-300
