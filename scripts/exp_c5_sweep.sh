#!/bin/bash
OUT=gpurun_out/${1:-c5s}; mkdir -p $OUT
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
for v in 15 17 24 25 26 8; do run v$v --variant $v; done
