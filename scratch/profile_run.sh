#!/bin/bash
# launch list of the timed step + ncu --set full of representative conv_halo launches (1 GPU)
set -x
CMD="python bench.py --steps 1 --warmup 3 --batch 4 --slots 2 --no-cpu-baseline"
$CMD > gpurun_out/plain_list.log 2>&1 || exit 1
L=$(tail -1 gpurun_out/plain_list.log | python -c "import sys,json; print(json.loads(sys.stdin.read())['gpu_launches'])")
echo "launches per step: $L"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L)) -c $L --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_list.log 2>&1
gzip -f gpurun_out/launches_r01b.csv
python scratch/conv_breakdown.py 8 5 > gpurun_out/plain_bd.log 2>&1 || exit 1
# conv_halo launches 40.. of the encoder at B=8: the first ResidualBottleneck blocks of g_a at 256x320
ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 2600 -c 14 -o gpurun_out/prof_halo python scratch/conv_breakdown.py 8 5 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
