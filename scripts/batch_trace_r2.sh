#!/bin/bash
mkdir -p gpurun_out
export CIAO_PROBE_BATCHES=4096 CIAO_SO=$PWD/ciaoalgorithms.jl_b200/libciao_cuda_prof.so
{
CIAO_BATCH_XPF=0 timeout 300 python scripts/batch_probe.py
CIAO_BATCH_STAGE_TABLE=0 timeout 300 python scripts/batch_probe.py
CIAO_BATCH_STAGES=1 timeout 300 python scripts/batch_probe.py
} > gpurun_out/batch_trace_r2.log 2>&1
cat gpurun_out/batch_trace_r2.log
