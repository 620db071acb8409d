#!/bin/bash
mkdir -p gpurun_out
(bash scripts/batch_sweep.sh 2>&1) > gpurun_out/r2h_batch_sweep.log
SAN_TIMEOUT=500 bash scripts/sanitize.sh > gpurun_out/r2h_sanitize_summary.log 2>&1
cat gpurun_out/r2h_batch_sweep.log gpurun_out/r2h_sanitize_summary.log
