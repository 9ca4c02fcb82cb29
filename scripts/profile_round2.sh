#!/bin/bash
# Round-2 ncu captures (run under gpurun, ONE GPU).  Outputs under gpurun_out/; scripts/ncu_extract.py turns them into profiles/.
#  1. launch list of the bench command (per-launch gpu__time_duration: cold-cache / serialised, kernel SHARES are what count)
#  2. --set full of every kernel of one warm step at the bench shape (stage, GEMM1, LayerNorm, GEMM2, recurrence, head)
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-latency --no-e2e --no-cpu --no-train --no-variants --no-rank4 --no-library > gpurun_out/r02_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"stage_features_16|gemm_tc2_kernel|layernorm_relu_16|gru_seq_kernel|gemm_tc_kernel" -s 6 -c 6 -f \
    -o gpurun_out/r02_prof_step python scripts/profile_forward.py 4096 64 > gpurun_out/r02_prof_step.log 2>&1
echo "full set rc=$?"
ncu -i gpurun_out/r02_prof_step.ncu-rep --page raw --csv > gpurun_out/r02_prof_step_raw.csv 2>/dev/null
ls -la gpurun_out/r02_prof_step* gpurun_out/r02_launches*
#  3. --set full of the persistent recurrence kernels of ONE training step (B 16 x T 128, tf32 mode): forward + BPTT
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gru_bptt2_kernel|gru_latency_kernel" -c 2 -f \
    -o gpurun_out/r02_prof_train python scripts/train_profile.py 16 tf32 > gpurun_out/r02_prof_train.log 2>&1
echo "training recurrence full set rc=$?"
ncu -i gpurun_out/r02_prof_train.ncu-rep --page raw --csv > gpurun_out/r02_prof_train_raw.csv 2>/dev/null
