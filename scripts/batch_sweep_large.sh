#!/bin/bash
# launch-shape sweep of the persistent minibatch kernel for LARGE batches (the rows phase dominates)
export CIAO_PROBE_BATCHES=65536,16384
for T in 64 128 256; do for C in 1 2 4 8; do
  CIAO_BATCH_T=$T CIAO_BATCH_CTAS=$C timeout 120 python scripts/batch_probe.py 2>&1 | grep "batch "
done; done
