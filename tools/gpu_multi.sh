#!/bin/bash
# Multi-GPU check (gpurun --gpus N): the product's multi-device entry points, then the bench under torchrun.  usage: tools/gpu_multi.sh <tag> <N>
tag=${1:-mg}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
( time timeout 600 python -m pytest tests/test_gpu_multi_device.py -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/${tag}_bench.err; cut -c1-400 gpurun_out/${tag}_bench.json
