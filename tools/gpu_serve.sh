#!/bin/bash
# batch command line on the GPU: its tests, then whole-process timings cold vs through the resident engine.  usage: tools/gpu_serve.sh <tag> [all]
tag=${1:-srv}
mkdir -p gpurun_out
sel='resident or bmp_qoi'; [ "$2" = "all" ] && sel='not nothing'
( timeout 600 python -m pytest tests/test_gpu_cli_batch.py tests/test_gpu_cli_dropin.py -x -q -s -k "$sel" ) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log; grep "resident engine" gpurun_out/${tag}_pytest.log
( timeout 300 python tools/cli_bench.py 64 jls; timeout 300 python tools/cli_bench.py 16 hevc ) > gpurun_out/${tag}_cli.log 2>&1; cat gpurun_out/${tag}_cli.log
