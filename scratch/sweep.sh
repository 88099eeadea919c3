#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "pipeline or graph" 2>&1 | tail -1
i=0
for cfg in "--threads 0" "--threads 1" "--threads 0" "--threads 1"; do
  i=$((i+1))
  python bench.py --no-cpu-baseline $cfg > gpurun_out/sweep_$i.log 2>&1
  echo "== $cfg (rc $?)"
  tail -1 gpurun_out/sweep_$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])" 2>/dev/null || tail -2 gpurun_out/sweep_$i.log | cut -c1-200
done
