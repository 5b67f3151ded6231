#!/bin/bash
# One gpurun call: GPU test tier, default bench (both arms), launch list.  Outputs under gpurun_out/<tag>_*.
# usage: tools/gpu_check.sh <tag> [tests|bench|launches ...]
tag=${1:-run}; shift
what=${*:-tests bench launches}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt
for w in $what; do
  case $w in
    tests)
      ( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 ) > gpurun_out/${tag}_pytest.log 2>&1
      tail -5 gpurun_out/${tag}_pytest.log ;;
    bench)
      ( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      echo "bench rc=$?"; head -c 1500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err ;;
    ref)
      ( time timeout 600 python bench.py --impl reference ) > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
      cat gpurun_out/${tag}_bench_ref.json ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip content,c5,parity > gpurun_out/${tag}_launches_bench.log 2>&1
      echo "launches rc=$?"; tail -3 gpurun_out/${tag}_launches.csv ;;
  esac
done
