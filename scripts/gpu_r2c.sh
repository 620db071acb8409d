#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hazard_stress.py -q 2>&1 | tail -40) > gpurun_out/r2c_multi_stress.log
(timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_hazard_stress.py 2>&1 | tail -40) > gpurun_out/r2c_rest.log
(timeout 600 python scripts/seq_variants.py --worker base 2>&1 | tail -40) > gpurun_out/r2c_variants.log
(timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2c_bench.err | tail -3) > gpurun_out/r2c_bench.json
tail -n 6 gpurun_out/r2c_*.log
