#!/bin/bash
# Round-2 evidence run on one B200: bench line, ncu launch list of the headline, ncu --set full of the kernels changed this round.
# Numbers printed under ncu are never bench values; the bench JSON comes from the first, un-profiled command.
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_single_gpu.json 2> gpurun_out/bench_r2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/ncu_launches_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:row_pass -s 3 -c 1 -f -o gpurun_out/prof_rowpass_r2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/ncu_rowpass_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pass_tail -s 3 -c 1 -f -o gpurun_out/prof_tail_r2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/ncu_tail_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:seq_kernel -s 1 -c 1 -f -o gpurun_out/prof_seq_r2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/ncu_seq_r2.log 2>&1
for k in "seq_kernel" "proshi_steps_kernel" "proshi_solution_kernel" "batch_persistent_kernel"; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -f -o gpurun_out/prof_cfg_${k}_r2 \
      python scripts/run_configs.py > gpurun_out/ncu_cfg_${k}_r2.log 2>&1
done
for f in gpurun_out/prof_*_r2.ncu-rep; do
  python scripts/ncu_summary.py $f gpurun_out/$(basename $f .ncu-rep).csv > /dev/null 2>&1
done
ls -la gpurun_out/*_r2.ncu-rep gpurun_out/*_r2.csv
