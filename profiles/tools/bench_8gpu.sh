#!/bin/bash
# The BASELINE configs that are quoted at 8 GPUs, one torchrun each (run under `gpurun --gpus 8`).
N=${1:-8}; OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) \
     bench.py --gpus $N "$@" > $OUT/r2q_${tag}_${N}gpu.json 2> $OUT/r2q_${tag}_${N}gpu.err; echo "$tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2q_${tag}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$tag", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), "batch", d["config"]["batch"], "pairs/step", d["config"]["pairs_per_step"], d["scaling"], (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print("$tag", e)
PY
}
run weak --steps 4 --warmup 3 --no-latency
run strong256 --steps 6 --warmup 3 --total-pairs 256 --no-cpu-baseline --no-latency
run r2d --steps 3 --warmup 3 --model ELIC_united_R2D --height 530 --width 730 --no-cpu-baseline --no-latency
run 1080p --steps 2 --warmup 3 --height 1080 --width 1920 --no-cpu-baseline --no-latency
