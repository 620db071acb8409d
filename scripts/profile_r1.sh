#!/bin/bash
# Round-1 evidence run on one B200: bench line, ncu launch list, ncu --set full of the two dominant kernels.
# Numbers printed under ncu are never bench values; the bench JSON comes from the first, un-profiled command.
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_single_gpu.json 2> gpurun_out/bench_r1.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:row_pass -s 3 -c 2 -f -o gpurun_out/prof_rowpass_r1 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rowpass.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:seq_kernel -s 1 -c 1 -f -o gpurun_out/prof_seq_r1 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_seq.log 2>&1
ls -la gpurun_out/*.ncu-rep
