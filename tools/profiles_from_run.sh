#!/bin/bash
# Copies the evidence of one tools/gpu_final.sh run from gpurun_out/ (scratch) into profiles/ (tracked), summarising the ncu reports
# here (no GPU needed).   usage: tools/profiles_from_run.sh <tag> <round>      e.g.  r2f r2
tag=$1; r=$2
cd "$(dirname "$0")/.."
for k in hevc_pair nearscan jls_pair; do
  rep=gpurun_out/${tag}_$k.ncu-rep
  [ -f $rep ] || continue
  name=$k; [ $k = nearscan ] && name=jls_nearscan
  ncu -i $rep --page details > profiles/ncu_${r}_${name}_details.txt 2>/dev/null
  ncu -i $rep --page source --print-source cuda,sass --csv > /tmp/${tag}_${k}_source.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/${tag}_${k}_source.csv 40 > profiles/ncu_${r}_${name}_lines.txt 2>&1
done
[ -f gpurun_out/${tag}_hevc_pair.ncu-rep ] && python tools/ncu_traffic.py gpurun_out/${tag}_hevc_pair.ncu-rep hevc_pair "round 2: 8 x 1920x1080 natural, qpd6 4, one launch of 16 CTAs in 8 clusters"
[ -f gpurun_out/${tag}_nearscan.ncu-rep ] && python tools/ncu_traffic.py gpurun_out/${tag}_nearscan.ncu-rep jls_nearscan "round 2: 32 x 3840x2160 RGB natural, NEAR 2, the device-resident launch of configs[3]'s per-GPU share: 96 planes"
[ -f gpurun_out/${tag}_jls_pair.ncu-rep ] && python tools/ncu_traffic.py gpurun_out/${tag}_jls_pair.ncu-rep jls_pair "round 2: 64 x 1920x1080 natural, NEAR 0, the device-resident step of bench.py, 128 CTAs in 64 clusters"
for f in bench.json bench_ref.json launches.csv pytest.log sanitizer.log cost.log cli.log; do
  [ -f gpurun_out/${tag}_$f ] || continue
  case $f in
    bench.json) cp gpurun_out/${tag}_$f profiles/bench_${r}.json ;;
    bench_ref.json) cp gpurun_out/${tag}_$f profiles/bench_${r}_ref.json ;;
    launches.csv) cp gpurun_out/${tag}_$f profiles/launches_${r}.csv ;;
    pytest.log) cp gpurun_out/${tag}_$f profiles/pytest_gpu_${r}.log ;;
    sanitizer.log) grep -v "^$" gpurun_out/${tag}_$f | cut -c1-220 > profiles/sanitizer_${r}.txt ;;
    cost.log) cp gpurun_out/${tag}_$f profiles/quick_timings_${r}.txt ;;
    cli.log) cp gpurun_out/${tag}_$f profiles/cli_bench_${r}.txt ;;
  esac
done
ls -la profiles | grep "_${r}[._]"
