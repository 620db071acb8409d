#!/bin/bash
mkdir -p gpurun_out
export CIAO_PROBE_BATCHES=4096 CIAO_SO=$PWD/ciaoalgorithms.jl_b200/libciao_cuda_prof.so
{
timeout 120 python scripts/batch_probe.py
} > gpurun_out/batch_trace_r2.log 2>&1
cat gpurun_out/batch_trace_r2.log
