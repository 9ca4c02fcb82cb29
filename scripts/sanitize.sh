#!/bin/bash
# compute-sanitizer over the kernels with hand-rolled cross-CTA protocols: the persistent batched
# recurrence (dataflow counters, CTA pairs), the CTA-pair GEMM (multicast commits, remote mbarrier arrives), the per-frame
# cooperative kernel (grid barrier, last-CTA tail) and the persistent BPTT kernel (tagged exchange words).
# Run on ONE GPU via gpurun:  bash scripts/sanitize.sh ; logs -> gpurun_out/r02_sanitizer_*.log
set -u
mkdir -p gpurun_out
T1="tests/test_gpu_parity.py::test_big_batch_tensor_recurrence_vs_oracle"
T2="tests/test_gpu_parity.py::test_gemm16_2cta"
T3="tests/test_gpu_parity.py::test_online_session_multi_stream_host_labels"
T4="tests/test_gpu_training.py::test_gradients_match_reference_golden"
run() {  # tool, name, timeout, pytest args...
    tool=$1; name=$2; to=$3; shift 3
    echo "=== $tool $name" | tee -a gpurun_out/r02_sanitizer_summary.txt
    timeout "$to" compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
        python -m pytest -x -q -m gpu "$@" > "gpurun_out/r02_sanitizer_${tool}_${name}.log" 2>&1
    rc=$?
    echo "rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' "gpurun_out/r02_sanitizer_${tool}_${name}.log" | tr '\n' ' ')" | tee -a gpurun_out/r02_sanitizer_summary.txt
}
: > gpurun_out/r02_sanitizer_summary.txt
for tool in memcheck racecheck; do
    run $tool recurrence 900 "$T1"
    run $tool gemm2cta 900 "$T2" -k "300 and fp16 and 3072"
    run $tool online 600 "$T3" -k "3"
    run $tool training 900 "$T4"
done
cat gpurun_out/r02_sanitizer_summary.txt
