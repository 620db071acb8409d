#!/bin/bash
# launch-shape sweep of the persistent minibatch kernel at C2 (threads per CTA, CTAs per SM, ring stages)
for T in 64 128; do for C in 2 4 8; do for S in 2 3; do
  CIAO_BATCH_T=$T CIAO_BATCH_CTAS=$C CIAO_BATCH_STAGES=$S timeout 120 python scripts/batch_probe.py 2>&1 | grep "batch 4096"
done; done; done
