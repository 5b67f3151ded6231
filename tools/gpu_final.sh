#!/bin/bash
# Round-end evidence in one gpurun call: full GPU tier, bench (both arms), launch list, ncu --set full of the three dominant
# kernels, sanitizers, the sweep cost model's inputs.   usage: tools/gpu_final.sh <tag> [steps...]
tag=${1:-fin}; shift
what=${*:-tests bench ref launches ncu sanitize cost cli}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for w in $what; do
  case $w in
    tests) ( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log ;;
    bench) ( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/${tag}_bench.err ;;
    ref) ( time timeout 600 python bench.py --impl reference ) > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; cut -c1-300 gpurun_out/${tag}_bench_ref.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip content,c5,parity > gpurun_out/${tag}_launches_bench.log 2>&1; echo "launches rc=$?" ;;
    ncu)
      timeout 600 $NCU -k regex:hevc_pair -s 1 -c 1 -f -o gpurun_out/${tag}_hevc_pair python tools/quick_hevc.py natural 8 4 1080 1920 2 2 > gpurun_out/${tag}_ncu_hevc.log 2>&1; tail -1 gpurun_out/${tag}_ncu_hevc.log
      timeout 600 $NCU -k regex:nearscan -s 1 -c 1 -f -o gpurun_out/${tag}_nearscan python tools/quick_c4.py 32 2 0 > gpurun_out/${tag}_ncu_near.log 2>&1; tail -1 gpurun_out/${tag}_ncu_near.log
      timeout 600 $NCU -k regex:jls_pair -s 2 -c 1 -f -o gpurun_out/${tag}_jls_pair python tools/quick_jls.py natural 64 0 > gpurun_out/${tag}_ncu_pair.log 2>&1; tail -1 gpurun_out/${tag}_ncu_pair.log ;;
    sanitize)
      ( timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py; timeout 900 compute-sanitizer --tool racecheck --racecheck-report hazard python tools/sanitize_small.py
        timeout 600 compute-sanitizer --tool racecheck --racecheck-report hazard python tools/race_hevc_tiny.py 64 64 ) > gpurun_out/${tag}_sanitizer.log 2>&1; grep -E "SUMMARY|ERROR|all paths|^ok" gpurun_out/${tag}_sanitizer.log ;;
    cost) ( for q in 1 2 3 4; do python tools/quick_hevc.py natural 16 $q 1080 1920 1 0; done; python tools/quick_hevc.py natural 96 4 1080 1920 1 0; python tools/quick_jls.py natural 64 2; python tools/quick_jls.py texture 64 0 ) > gpurun_out/${tag}_cost.log 2>&1; cat gpurun_out/${tag}_cost.log ;;
    cli) ( python tools/cli_bench.py 64 jls; python tools/cli_bench.py 16 hevc ) > gpurun_out/${tag}_cli.log 2>&1; tail -6 gpurun_out/${tag}_cli.log ;;
  esac
done
