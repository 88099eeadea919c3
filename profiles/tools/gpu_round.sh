#!/bin/bash
# One GPU-box session: tests, smoke, bench (both arms), ncu launch list + DRAM traffic, compute-sanitizer.
# Usage (from the repo root on the box): bash profiles/tools/gpu_round.sh <tag> [steps...]
TAG=${1:-r02}; shift
STEPS=${@:-tests smoke bench ref ncu sanitize}
OUT=gpurun_out; mkdir -p $OUT
for s in $STEPS; do case $s in
tests)   : > $OUT/${TAG}_tests.log
         for f in tests/test_gpu_*.py; do   # one process per file: a device-side trap in one file must not take the others down
           timeout 1500 python -m pytest $f -m gpu -q -s >> $OUT/${TAG}_tests.log 2>&1; echo "$f rc=$?"
         done
         grep -E "passed|failed|error" $OUT/${TAG}_tests.log | tail -12; grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests.log | head -20; grep "\[parity" $OUT/${TAG}_tests.log;;
smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log;;
bench)   timeout 900 python bench.py --steps 4 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -c 4000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err;;
ref)     timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>&1; echo "ref rc=$?"; tail -c 1500 $OUT/${TAG}_bench_ref.json;;
ncu)     timeout 900 python profiles/tools/plan_once.py ELIC_united 480 640 8 bf16 > $OUT/${TAG}_plan_once.log 2>&1 && \
         timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
            --clock-control none --csv --log-file $OUT/${TAG}_plan.csv python profiles/tools/plan_once.py ELIC_united 480 640 8 bf16 > $OUT/${TAG}_plan_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/${TAG}_plan_once.log; wc -l $OUT/${TAG}_plan.csv;;
sanitize) for tool in memcheck racecheck; do
           timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/plan_once.py ELIC_united 128 128 1 bf16 mid > $OUT/${TAG}_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|PLAN" $OUT/${TAG}_sanitizer_$tool.log | tail -3
         done;;
esac; done
