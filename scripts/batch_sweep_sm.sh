#!/bin/bash
# launch-shape knobs of batch_sm_kernel at C2 (sub-groups, threads per sub-group, ring depth)
mkdir -p gpurun_out
export CIAO_PROBE_BATCHES=4096,65536
{
  timeout 200 python scripts/batch_probe.py
  CIAO_BATCH_T=256 timeout 200 python scripts/batch_probe.py
  CIAO_BATCH_CTAS=2 timeout 200 python scripts/batch_probe.py
  CIAO_BATCH_STAGES=3 timeout 200 python scripts/batch_probe.py
  CIAO_BATCH_STAGES=1 timeout 200 python scripts/batch_probe.py
} > gpurun_out/batch_sweep_sm.log 2>&1
cat gpurun_out/batch_sweep_sm.log
