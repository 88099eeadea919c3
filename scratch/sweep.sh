#!/bin/bash
i=0
while read -r cfg; do
  i=$((i+1))
  python bench.py --no-cpu-baseline $cfg > gpurun_out/sweep_$i.log 2>&1
  echo "== $cfg (rc $?)"
  tail -1 gpurun_out/sweep_$i.log | cut -c1-160
done <<'CFGS'
--slots 8 --batch 16 --steps 4
--slots 6 --batch 16 --steps 4
CFGS
