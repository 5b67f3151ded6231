#!/bin/bash
# HEVC iteration check on the GPU: parity tests, racecheck on a few CTUs, timings.  usage: tools/gpu_hevc_iter.sh <tag> [fast]
tag=${1:-hv}
mkdir -p gpurun_out
if [ "$2" != "fast" ]; then
( time timeout 900 python -m pytest tests/test_gpu_hevc.py tests/test_gpu_random.py -x -q -k "not 4k_qp_sweep" ) > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
( timeout 600 compute-sanitizer --tool racecheck --racecheck-report hazard python tools/race_hevc_tiny.py 32 64 ) > gpurun_out/${tag}_race.log 2>&1; tail -2 gpurun_out/${tag}_race.log
else
( time timeout 900 python -m pytest tests/test_gpu_hevc.py -x -q -k "golden or edge" ) > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
fi
( python tools/quick_hevc.py natural 64 4 1080 1920 2 0; python tools/quick_hevc.py texture 64 4 1080 1920 1 0; python tools/quick_hevc.py natural 96 4 1080 1920 1 0
  python tools/quick_hevc.py natural 16 4 1080 1920 1 0; python tools/quick_hevc.py natural 16 1 1080 1920 1 0 ) > gpurun_out/${tag}_quick.log 2>&1; cat gpurun_out/${tag}_quick.log
( IMCVT_B200_LIB=libimcvt_b200_prof.so python tools/prof_hevc.py natural 16 4 1080 1920 2 ) > gpurun_out/${tag}_prof.log 2>&1; cat gpurun_out/${tag}_prof.log
