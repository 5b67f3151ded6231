#!/bin/bash
# resident engine on the GPU: its test, then whole-process timings cold vs through the server.  usage: tools/gpu_serve.sh <tag>
tag=${1:-srv}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_cli_batch.py -x -q -s -k "resident or bmp_qoi" ) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log; grep "resident engine" gpurun_out/${tag}_pytest.log
( timeout 300 python tools/cli_bench.py 64 jls; timeout 300 python tools/cli_bench.py 16 hevc ) > gpurun_out/${tag}_cli.log 2>&1; cat gpurun_out/${tag}_cli.log
