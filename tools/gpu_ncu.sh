#!/bin/bash
# One gpurun call: ncu --set full captures of the dominant kernels + phase profiles.  usage: tools/gpu_ncu.sh <tag> [hevc|near|pair|prof|quick ...]
tag=${1:-ncu}; shift
what=${*:-quick prof hevc near pair}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for w in $what; do
  case $w in
    quick)
      ( python tools/quick_jls.py natural 64 2; python tools/quick_jls.py texture 64 2; python tools/quick_jls.py natural 64 0
        python tools/quick_hevc.py natural 64 4 1080 1920 2 0; python tools/quick_hevc.py natural 64 4 1080 1920 2 10
        python tools/quick_hevc.py natural 96 4 1080 1920 2 0; python tools/quick_hevc.py natural 96 4 1080 1920 2 5 ) > gpurun_out/${tag}_quick.log 2>&1
      cat gpurun_out/${tag}_quick.log ;;
    prof)
      ( python tools/prof_hevc.py natural 16 4 1080 1920 2; python tools/prof_hevc.py natural 16 4 1080 1920 10
        python tools/prof_hevc.py texture 16 4 1080 1920 2 ) > gpurun_out/${tag}_prof_hevc.log 2>&1
      cat gpurun_out/${tag}_prof_hevc.log ;;
    hevc)
      timeout 900 $NCU -k regex:hevc_pair -s 1 -c 1 -f -o gpurun_out/${tag}_hevc_pair python tools/quick_hevc.py natural 8 4 1080 1920 2 2 > gpurun_out/${tag}_hevc_ncu.log 2>&1; tail -2 gpurun_out/${tag}_hevc_ncu.log ;;
    near)
      timeout 600 $NCU -k regex:nearscan -s 1 -c 1 -f -o gpurun_out/${tag}_nearscan python tools/quick_c4.py 2 2 1 > gpurun_out/${tag}_near_ncu.log 2>&1; tail -2 gpurun_out/${tag}_near_ncu.log ;;
    pair)
      timeout 600 $NCU -k regex:jls_pair -s 2 -c 1 -f -o gpurun_out/${tag}_jls_pair python tools/quick_jls.py natural 64 0 > gpurun_out/${tag}_pair_ncu.log 2>&1; tail -2 gpurun_out/${tag}_pair_ncu.log ;;
  esac
done
ls -la gpurun_out
