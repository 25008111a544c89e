#!/bin/bash
# 2 GPUs: the landmark-sharded config-5 solve on the stream (KBA_SHARD_GRAPH=0) and as a captured graph with the NCCL
# all-reduces inside (1), each checked against the one-GPU solve; then the bench line on 2 ranks
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
KBA_SHARD_GRAPH=0 timeout 150 $TR --master-port 29511 scripts/config5_sharded.py --steps 3 --check 2>&1 | grep -v Warning | tail -2 | tee gpurun_out/r2_shard23_stream.log
KBA_SHARD_GRAPH=1 KBA_GRAPH_VERBOSE=1 timeout 150 $TR --master-port 29512 scripts/config5_sharded.py --steps 3 --check 2>&1 | grep -v Warning | tail -4 | tee gpurun_out/r2_shard23_graph.log
G=1; grep -q '"ms_per_solve"' gpurun_out/r2_shard23_graph.log || G=0
echo "shard graph usable: $G"
KBA_SHARD_GRAPH=$G timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/r2_bench23_2gpu.json 2> gpurun_out/r2_bench23_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench23_2gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench23_2gpu.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
print(json.dumps(d["sub_records"]["config5"])[:1200])
PY
