#!/bin/bash
# 1/2/4/8-GPU scaling of the sharded full-gradient pass (weak: 2^22 rows per GPU, C4-style) and one
# strong-scaling SVRG++ point at 8 GPUs.  Run on an 8-GPU box:  gpurun --gpus 8 -- bash scripts/scale8.sh
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    python bench.py --workload fullgrad --steps 10 --no-cpu-baseline > gpurun_out/scale_fullgrad_n$n.json 2> gpurun_out/scale_fullgrad_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --workload fullgrad --steps 10 > gpurun_out/scale_fullgrad_n$n.json 2> gpurun_out/scale_fullgrad_n$n.err
  fi
  python - <<PY
import json
j=json.loads(open("gpurun_out/scale_fullgrad_n$n.json").read().strip().splitlines()[-1])
print("fullgrad n=$n value", round(j["value"],3), "epochs/s; aggregate GB/s", round(j["full_gradient"]["aggregate_gbs"],1), "per-GPU kernel GB/s", round(j["full_gradient"]["gbs_per_gpu"],1), "e2e", round(j["e2e"]["value"],3))
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/scale_svrgpp_n8.json 2> gpurun_out/scale_svrgpp_n8.err
python - <<PY
import json
j=json.loads(open("gpurun_out/scale_svrgpp_n8.json").read().strip().splitlines()[-1])
print("svrgpp n=8 value", round(j["value"],3), "pass ms", j["svrg"]["pass_ms"], "us/step", round(j["svrg"]["us_per_inner_step"],3))
PY
