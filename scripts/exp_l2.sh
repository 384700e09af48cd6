#!/bin/bash
# flux reductions with evict_last priority + L2 fetch granularity
OUT=gpurun_out/${1:-l2}; mkdir -p $OUT
run() { local name=$1; shift
  timeout 900 python bench.py --no-cpu --no-e2e "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  echo "$name rc=$? $(python - "$OUT/bench_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print("value=%.3e ms/step=%.3f frac=%.3f variant=%s"%(d["value"],d["ms_per_step"],r["frac"],d["config"].get("variant")))
except Exception as e: print("parse-fail",e)
PY
)" | tee -a "$OUT/summary.txt"; }
python - <<'PY'
from pumiumtally_b200.tally import PumiTally
e=PumiTally.from_spec("box:2,2,2",10); print("default l2_fetch", e.get_option("l2_fetch"))
PY
run c2_v8 --steps 10 --warmup 3 --variant 8
run c2_v8_f32 --steps 10 --warmup 3 --variant 8 --opt l2_fetch=32
run c2_v8_f64 --steps 10 --warmup 3 --variant 8 --opt l2_fetch=64
run c2_v8_f128 --steps 10 --warmup 3 --variant 8 --opt l2_fetch=128
run c2_v20 --steps 5 --warmup 3 --variant 20
run c3_v8 --config c3 --steps 3 --warmup 3 --variant 8
run c5_v16 --config c5 --per-gpu-share --steps 3 --warmup 3 --variant 16
run c5_v16_f32 --config c5 --per-gpu-share --steps 3 --warmup 3 --variant 16 --opt l2_fetch=32
