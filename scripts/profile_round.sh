#!/bin/bash
# Round-end ncu captures (run under gpurun, ONE GPU).  Outputs under gpurun_out/; summaries are copied to profiles/ by hand.
#  1. launch list of the bench command (per-launch gpu__time_duration, cold-cache/serialised: kernel SHARES are what count)
#  2. --set full of the projection GEMMs at the bench shape (one GEMM1 + one GEMM2 launch of a warm step)
#  3. --set full of the fused per-frame kernel, caches left warm (its weights live in L2 between frames)
#  4. --set full of the persistent BPTT kernel (training step)
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum"
ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv \
    python bench.py --steps 2 --warmup 3 --no-latency --no-e2e --no-cpu --no-train --no-variants > gpurun_out/launches_r01b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 6 -c 2 -f -o gpurun_out/prof_r01_gemm \
    python scripts/profile_forward.py 4096 64 > gpurun_out/prof_r01_gemm.log 2>&1
ncu -i gpurun_out/prof_r01_gemm.ncu-rep --page raw --csv > gpurun_out/raw_r01_gemm.csv 2>/dev/null
ncu --set full --clock-control none --cache-control none --import-source on --graph-profiling node -k regex:online_fused -s 60 -c 2 -f \
    -o gpurun_out/prof_r01_online python scripts/online_latency.py > gpurun_out/prof_r01_online.log 2>&1
ncu -i gpurun_out/prof_r01_online.ncu-rep --page raw --csv > gpurun_out/raw_r01_online.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:gru_bptt_kernel\|gru_latency_kernel -s 16 -c 2 -f \
    -o gpurun_out/prof_r01_train python scripts/train_profile.py 16 > gpurun_out/prof_r01_train.log 2>&1
ncu -i gpurun_out/prof_r01_train.ncu-rep --page raw --csv > gpurun_out/raw_r01_train.csv 2>/dev/null
ls -la gpurun_out/*r01* | tail -20
