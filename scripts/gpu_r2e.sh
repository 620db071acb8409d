#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hazard_stress.py -q 2>&1 | tail -40) > gpurun_out/r2e_multi_stress.log
(timeout 600 python scripts/seq_variants.py --worker base 2>&1 | tail -40) > gpurun_out/r2e_variants.log
(timeout 300 python scripts/prof_seq.py 20 1024 saga,finito ls x 2>&1 | tail) > gpurun_out/r2e_prof.log
(timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_hazard_stress.py 2>&1 | tail -40) > gpurun_out/r2e_rest.log
tail -n 12 gpurun_out/r2e_*.log
