#!/bin/bash
# Runs the GPU parity suite in isolated chunks (each under its own timeout so a hung kernel
# cannot take the whole call down) and collects logs under gpurun_out/.
# Usage (via gpurun):  bash scripts/gpu_ci.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
run() {  # name, timeout, pytest -k expr
    echo "=== $1" | tee -a gpurun_out/ci_summary.txt
    timeout "$2" python -m pytest tests -m gpu -q -x -s -k "$3" > "gpurun_out/ci_$1.log" 2>&1
    rc=$?
    echo "rc=$rc $(tail -n 1 gpurun_out/ci_$1.log)" | tee -a gpurun_out/ci_summary.txt
}
: > gpurun_out/ci_summary.txt
run gemm_tc 300 "gemm16_tcgen05"
run gemm_2cta 300 "gemm16_2cta or gemm_tf32"
run simt 300 "gemm_f32 or forward_fp32"
run aggregate 300 "aggregate"
run bf16 400 "forward_bf16 or forward_fp16 or pooled or online or chunking or carried or single_frame or big_batch or module_forward or evaluate or batched_ragged"
run fullsize 400 "full_size"
run training 400 "gradients or dropout_mask"
cat gpurun_out/ci_summary.txt
