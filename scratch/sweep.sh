#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "conv_tc" 2>&1 | tail -1
for v in 96 0; do
  RGBD_TC_SMALLRING=$v python bench.py --no-cpu-baseline > gpurun_out/sweep_ring$v.log 2>&1
  echo "== smallring $v (rc $?)"
  tail -1 gpurun_out/sweep_ring$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"
done
