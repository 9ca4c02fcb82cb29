#!/bin/bash
# ncu capture of the fused per-frame kernel (run under gpurun): full set on 2 launches after warm-up.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:online_fused -s 40 -c 2 \
    -o gpurun_out/prof_online -f python scripts/online_latency.py > gpurun_out/prof_online.log 2>&1
ncu -i gpurun_out/prof_online.ncu-rep --page raw --csv > gpurun_out/raw_online.csv 2>/dev/null
ncu -i gpurun_out/prof_online.ncu-rep --page source --csv > gpurun_out/src_online.csv 2>/dev/null
tail -5 gpurun_out/prof_online.log
