#!/bin/bash
# 2 GPUs: independent windows per rank + the landmark-sharded config-5 solve (folded exchanges) checked against the one-GPU solve
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/r2_bench16_2gpu.json 2> gpurun_out/r2_bench16_2gpu.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench16_2gpu.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench16_2gpu.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
print(json.dumps(d["sub_records"], indent=1)[:3000])
PY
