#!/bin/bash
# launch list of the timed step + ncu --set full of representative conv_halo launches (1 GPU)
set -x
CMD="python bench.py --steps 1 --warmup 3 --batch 4 --slots 2 --graphs 0 --no-cpu-baseline"   # eager launches: the lib counter == the ncu launch index
$CMD > gpurun_out/plain_list.log 2>&1 || exit 1
L=$(tail -1 gpurun_out/plain_list.log | python -c "import sys,json; print(json.loads(sys.stdin.read())['gpu_launches'])")
echo "launches per step: $L"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L)) -c $L --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_list.log 2>&1
gzip -f gpurun_out/launches_r01b.csv
# ncu --set full of single-layer launches (one launch per layer, no tracing)
RGBD_NCU=1 python scratch/tc_trace.py 8 > gpurun_out/plain_tr.log 2>&1 || exit 1
RGBD_NCU=1 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -o gpurun_out/prof_halo python scratch/tc_trace.py 8 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
