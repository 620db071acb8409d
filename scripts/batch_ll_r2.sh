#!/bin/bash
# one-CTA-per-SM flagged-word exchange vs grid barriers in the persistent minibatch kernel (C2 shape), plus the minibatch parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_minibatch.py -x -q > gpurun_out/pytest_minibatch_ll.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_minibatch_ll.log
tail -5 gpurun_out/pytest_minibatch_ll.log
export CIAO_PROBE_BATCHES=4096,512,16384,65536
{
  timeout 300 python scripts/batch_probe.py
  CIAO_BATCH_TWO_DOTS=1 timeout 300 python scripts/batch_probe.py
  CIAO_BATCH_EXCHANGE=barrier timeout 300 python scripts/batch_probe.py
  CIAO_BATCH_STAGES=3 timeout 300 python scripts/batch_probe.py
  CIAO_PROBE_BATCHES=4096 CIAO_SO=$PWD/ciaoalgorithms.jl_b200/libciao_cuda_prof.so timeout 300 python scripts/batch_probe.py
} > gpurun_out/batch_ll_r2.log 2>&1
cat gpurun_out/batch_ll_r2.log
