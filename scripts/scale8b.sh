#!/bin/bash
# Reduced 8-GPU evidence run (round 1, after the kernel work of session 2): weak-scaling end points of the sharded
# full-gradient pass, the strong-scaling SVRG++ point at 8 GPUs, and the C4-size row-sharded solve on 4 GPUs.
#   gpurun --gpus 8 -- bash scripts/scale8b.sh
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
python bench.py --workload fullgrad --steps 10 --no-cpu-baseline > gpurun_out/scale_fullgrad_n1.json 2> gpurun_out/scale_fullgrad_n1.err
run 8 29608 --workload fullgrad --steps 10 > gpurun_out/scale_fullgrad_n8.json 2> gpurun_out/scale_fullgrad_n8.err
run 8 29650 --steps 5 --warmup 3 > gpurun_out/scale_svrgpp_n8.json 2> gpurun_out/scale_svrgpp_n8.err
run 4 29660 --workload svrgpp-sharded --steps 2 --warmup 3 > gpurun_out/svrgpp_sharded_n4_C4.json 2> gpurun_out/svrgpp_sharded_n4_C4.err
python - <<PY
import json
for f in ("scale_fullgrad_n1", "scale_fullgrad_n8", "scale_svrgpp_n8", "svrgpp_sharded_n4_C4"):
    try:
        j = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(j["value"], 3), j["unit"], "ms/step", round(j["ms_per_step"], 3), "pass GB/s/GPU", round(j["full_gradient"]["gbs_per_gpu"], 1),
              "inner us/step", j.get("svrg", {}).get("us_per_inner_step"))
    except Exception as ex:
        print(f, "FAILED", ex, open(f"gpurun_out/{f}.err").read()[-600:])
PY
