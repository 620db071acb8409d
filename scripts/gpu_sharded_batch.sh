#!/bin/bash
# usage: gpu_sharded_batch.sh G [tests]   — sharded minibatch probe on G GPUs (and the multi-GPU tests when a second argument is given)
mkdir -p gpurun_out
G=${1:-2}
if [ -n "$2" ]; then (timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5) > gpurun_out/sharded_multi_tests_n$G.log; cat gpurun_out/sharded_multi_tests_n$G.log; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29577 scripts/sharded_batch_probe.py > gpurun_out/sharded_batch_n$G.log 2>&1
grep "GPUs\]\|Error\|error" gpurun_out/sharded_batch_n$G.log | head -30
