#!/bin/bash
# ncu --set full captures (one launch each) of the C2 / C5 / adaptive kernels exercised by scripts/run_configs.py.
mkdir -p gpurun_out
for k in "seq_kernel" "proshi_steps_kernel" "proshi_batch_kernel" "batch_persistent_kernel" "adaptive_kernel" "row_pass_kernel<8, 1"; do
  tag=$(echo "$k" | tr -c 'a-z_0-9' '_' | sed 's/__*/_/g; s/_$//')
  ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -f -o gpurun_out/prof_cfg_$tag \
      python scripts/run_configs.py > gpurun_out/ncu_cfg_$tag.log 2>&1
done
ls -la gpurun_out/prof_cfg_*.ncu-rep
