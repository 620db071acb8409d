#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_minibatch.py tests/test_gpu_parity.py -q -x -k "minibatch or full_gradient or saga_steps or finito_steps or lfinito" 2>&1 | tail -8) > gpurun_out/r2j_tests.log
for pf in 0 1; do
  CIAO_BATCH_L2PF=$pf python scripts/batch_probe.py 2>&1 | sed "s/^/[l2pf=$pf] /"
  CIAO_BATCH_L2PF=$pf CIAO_SO=$PWD/ciaoalgorithms.jl_b200/libciao_cuda_prof.so python scripts/batch_probe.py 2>&1 | grep cycles | sed "s/^/[l2pf=$pf] /"
done > gpurun_out/r2j_batch.log 2>&1
python scripts/k2_probe.py >> gpurun_out/r2j_batch.log 2>&1
cat gpurun_out/r2j_tests.log gpurun_out/r2j_batch.log
