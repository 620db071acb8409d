#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 5 --warmup 3 2> gpurun_out/r2n4_bench.err | tail -1) > gpurun_out/r2n4_bench.json
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2n4_bench.json').read().strip().splitlines()[-1])
print({k:b[k] for k in ('value','n_gpus','ms_per_step')})
print(b.get('c2_sharded_minibatch'))
print({k:v for k,v in b['c4']['full_gradient'].items() if k in ('ms_per_pass','tail_us','bitwise_equal_across_ranks')})
PY
tail -n 5 gpurun_out/r2n4_bench.err
