#!/bin/bash
# ncu --set full (with source) of the one-CTA-per-SM minibatch kernel at C2, batch 4096: second Finito launch and second LFinito launch of the probe
set -x
mkdir -p gpurun_out
CIAO_PROBE_BATCHES=4096 python scripts/batch_probe.py > gpurun_out/batch_sm_probe.log 2>&1 || exit 1
CIAO_PROBE_BATCHES=4096 ncu --set full --clock-control none --import-source on -k regex:batch_sm -s 1 -c 1 -f \
    -o gpurun_out/prof_finito_batch_sm_r2 python scripts/batch_probe.py > gpurun_out/ncu_fin_sm.log 2>&1
CIAO_PROBE_BATCHES=4096 ncu --set full --clock-control none --import-source on -k regex:batch_sm -s 3 -c 1 -f \
    -o gpurun_out/prof_lfinito_batch_sm_r2 python scripts/batch_probe.py > gpurun_out/ncu_lfin_sm.log 2>&1
ls -la gpurun_out/*.ncu-rep
