#!/bin/bash
i=0
while read -r cfg; do
  i=$((i+1))
  python bench.py --no-cpu-baseline $cfg > gpurun_out/sweep_$i.log 2>&1
  echo "== $cfg (rc $?)"
  tail -1 gpurun_out/sweep_$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['config']['hbm_peak_gb'])" 2>/dev/null || tail -2 gpurun_out/sweep_$i.log | cut -c1-200
done <<'CFGS'
--slots 8 --dec-slots 8
--slots 5 --dec-slots 11
--slots 6 --dec-slots 12
--slots 8 --dec-slots 8
--slots 5 --dec-slots 11
CFGS
