#!/bin/bash
# ncu --set full (with source) of the LFinito persistent minibatch kernel at C2, batch 4096: fourth launch of the probe
set -x
mkdir -p gpurun_out
CIAO_PROBE_BATCHES=4096 python scripts/batch_probe.py > gpurun_out/lfin_probe.log 2>&1 || exit 1
CIAO_PROBE_BATCHES=4096 ncu --set full --clock-control none --import-source on -k regex:batch_persistent -s 3 -c 1 -f \
    -o gpurun_out/prof_lfinito_batch_r2 python scripts/batch_probe.py > gpurun_out/ncu_lfin.log 2>&1
ls -la gpurun_out/
