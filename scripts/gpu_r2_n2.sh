#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
(timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -30) > gpurun_out/r2n2_multi.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/r2n2_bench.err | tail -3) > gpurun_out/r2n2_bench.json
tail -n 20 gpurun_out/r2n2_multi.log; tail -c 3000 gpurun_out/r2n2_bench.json; tail -n 20 gpurun_out/r2n2_bench.err
