#!/bin/bash
# the driver's round-end sequence on one B200: gpu tests, smoke, default bench (both arms)
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/final_pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/final_smoke.log
(timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1) > gpurun_out/final_bench_reference.json
(timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/final_bench.err | tail -1) > gpurun_out/final_bench.json
tail -n 6 gpurun_out/final_pytest_gpu.log gpurun_out/final_smoke.log; tail -c 600 gpurun_out/final_bench_reference.json; echo; tail -c 400 gpurun_out/final_bench.json
