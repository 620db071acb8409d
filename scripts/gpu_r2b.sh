#!/bin/bash
mkdir -p gpurun_out
python -c "import os,psutil;print('host cores',os.cpu_count(),'mem GiB',psutil.virtual_memory().total>>30)" > gpurun_out/r2b_host.log 2>&1
(timeout 600 python scripts/proshi_race_probe.py 2>&1 | tail -80) > gpurun_out/r2b_proshi_probe.log
(timeout 900 python scripts/seq_variants.py 2>&1 | tail -120) > gpurun_out/r2b_variants.log
(timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2b_bench.err | tail -3) > gpurun_out/r2b_bench.json
tail -n 5 gpurun_out/r2b_*.log
