mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|Thread|Core|NUMA node\(s\)"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_rounded or 16bit_features" > gpurun_out/hostround_tests.log 2>&1; echo rc=$?; tail -3 gpurun_out/hostround_tests.log
timeout 900 python bench.py --no-train --no-latency --no-rank4 --no-variants > gpurun_out/bench_v19.json 2> gpurun_out/bench_v19.err; echo bench rc=$?; tail -3 gpurun_out/bench_v19.err
python -c "
import json; d=json.load(open('gpurun_out/bench_v19.json')); print(json.dumps(d['e2e'],indent=1)); print(d['value'], d['cpu_baseline'])"
