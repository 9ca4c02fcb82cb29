#!/bin/bash
# Staging-overlap sweep: mode (0 none, 1 under GEMM1, 2 under the recurrence) x internal sub-chunk.
mkdir -p gpurun_out
: > gpurun_out/overlap_sweep.txt
for sub in 64 32 16; do
  for mode in 0 1 2; do
    if [ $sub = 64 ] && [ $mode != 0 ]; then continue; fi
    PREGO_STAGE_OVERLAP=$mode python bench.py --no-latency --no-e2e --no-cpu --no-train --no-variants --subchunk $sub --steps 8 --warmup 3 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sub=$sub mode=$mode', round(d['value']/1e6,2), 'Mfps', d['ms_per_step'], d['roofline']['phase_share'], d['clocks'])" >> gpurun_out/overlap_sweep.txt
  done
done
cat gpurun_out/overlap_sweep.txt
