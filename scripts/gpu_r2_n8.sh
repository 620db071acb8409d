#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/r2n8_bench.err | tail -3) > gpurun_out/r2n8_bench.json
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 2> gpurun_out/r2n4_bench.err | tail -3) > gpurun_out/r2n4_bench.json
tail -c 1500 gpurun_out/r2n8_bench.json; tail -n 5 gpurun_out/r2n8_bench.err
