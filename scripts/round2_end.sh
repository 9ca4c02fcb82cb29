#!/bin/bash
# Round-end verification on one B200 (run via gpurun): the GPU suite as the driver runs it, smoke(), the default bench, the reference
# arm, and the ncu captures (launch list + --set full of the step kernels).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_suite_final.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/r02_gpu_suite_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"; head -c 300 gpurun_out/r02_bench_reference.json; echo
bash scripts/profile_round2.sh
python - <<'P'
import json
d = json.load(open("gpurun_out/r02_bench_final.json"))
print("value", d["value"], "sustained", d["sustained"]["value"], "e2e", d["e2e"]["value"], d["e2e"].get("path"), "cpu", d["cpu_baseline"]["value"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "clocks", d["clocks"])
P
