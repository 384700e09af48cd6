#!/bin/bash
# packed binned variant (24) vs defaults
OUT=gpurun_out/${1:-packed}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "24 or 25" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
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
run c2_v24 --steps 10 --warmup 3 --variant 24
run c2_v25 --steps 10 --warmup 3 --variant 25
run c2_v26 --steps 10 --warmup 3 --variant 26
run c5_v25 --config c5 --per-gpu-share --steps 3 --warmup 3 --variant 25
run c5_v26 --config c5 --per-gpu-share --steps 3 --warmup 3 --variant 26
run c4_v25 --config c4 --steps 5 --warmup 3 --variant 25
run c3_v24 --config c3 --steps 3 --warmup 3 --variant 24
