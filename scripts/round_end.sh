#!/bin/bash
# Round-end verification on one B200 (run via gpurun): the GPU suite as the driver runs it, smoke(), the default bench,
# and the launch list of a short bench run (ncu; per-launch times are cold-cache / serialised: compare SHARES).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/full_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/full_gpu_suite.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; head -c 400 gpurun_out/bench_reference.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 4 --warmup 3 --no-latency --no-e2e --no-cpu --no-train --no-variants --no-rank4 > gpurun_out/launches_final.log 2>&1; echo "launch list rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_final.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("path"), "cpu", d["cpu_baseline"]["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"])
P
