#!/bin/bash
# Launch list + one full capture of the per-frame mAP kernels (run via gpurun):  bash scripts/ncu_ap.sh [N] [K]
N=${1:-1048576}; K=${2:-86}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ap_ -s 32 -c 16 --csv --log-file gpurun_out/ap_launches.csv python scripts/ap_bench.py $N $K > gpurun_out/ap_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ap_scatter -s 8 -c 4 -o gpurun_out/ap_scatter -f python scripts/ap_bench.py $N $K > gpurun_out/ap_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ap_hist|ap_scan_final|ap_build' -s 12 -c 4 -o gpurun_out/ap_other -f python scripts/ap_bench.py $N $K >> gpurun_out/ap_ncu_full.log 2>&1
grep -v "^==" gpurun_out/ap_launches.csv | awk -F'","' '{print $5, $NF}' | tail -18
