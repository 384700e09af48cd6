#!/bin/bash
OUT=gpurun_out/${1:-c5m}; mkdir -p $OUT
run() { local name=$1; shift
  timeout 900 python bench.py --no-cpu --no-e2e --config c5 --per-gpu-share --steps 3 --warmup 3 "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  echo "$name rc=$? $(python - "$OUT/bench_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print("value=%.3e ms/step=%.3f frac=%.3f variant=%s"%(d["value"],d["ms_per_step"],r["frac"],d["config"].get("variant")))
except Exception as e: print("parse-fail",e)
PY
)" | tee -a "$OUT/summary.txt"; }
run zmajor
export PUMITALLY_TET_ORDER=morton
run tetmorton_binzmajor
run tetmorton_binmorton --opt morton=1
run tetmorton_binmorton_c1 --opt morton=1 --opt claim_run=1
run tetmorton_binmorton_c2 --opt morton=1 --opt claim_run=2
run tetmorton_v24 --opt morton=1 --variant 24
