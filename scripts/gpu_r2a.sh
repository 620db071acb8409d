#!/bin/bash
# round-2 session A: new tail kernel / peer exchange / proxy fences on the GPU
mkdir -p gpurun_out
nvidia-smi -L | head -3
(timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -40) > gpurun_out/r2a_multi.log
(timeout 600 python -m pytest tests/test_gpu_hazard_stress.py -q 2>&1 | tail -40) > gpurun_out/r2a_stress.log
(timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_hazard_stress.py 2>&1 | tail -40) > gpurun_out/r2a_rest.log
(timeout 300 python scripts/exchange_floor.py 2>&1 | tail -60) > gpurun_out/r2a_floor.log
tail -5 gpurun_out/r2a_*.log
