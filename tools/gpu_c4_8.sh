#!/bin/bash
# configs[3] at its stated size: 256 x 3840x2160 RGB, NEAR 2, frame-sharded over 8 GPUs (gpurun --gpus 8).  usage: tools/gpu_c4_8.sh <tag> <N> [skip-list]
tag=${1:-c4x8}; N=${2:-8}; skip=${3:-content,hevc,c5}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt; nproc >> gpurun_out/${tag}_gpus.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --skip $skip ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -4 gpurun_out/${tag}_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().split("\n")[-1])
print("main", d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("parity"))
c=d.get("c4")
if c: print("c4", c["value"], c["e2e"]["value"], c["ms_per_step"], c["config"]["share"], c.get("parity"))
PY
