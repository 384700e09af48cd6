#!/bin/bash
OUT=gpurun_out/${1:-fin}; mkdir -p $OUT
timeout 1400 python -m pytest tests -m gpu -q -x --timeout=600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
run() { local name=$1; shift
  timeout 900 python bench.py --no-cpu --no-e2e "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  echo "$name rc=$? $(python - "$OUT/bench_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print("value=%.3e ms/step=%.3f frac=%.3f variant=%s launches=%s"%(d["value"],d["ms_per_step"],r["frac"],d["config"].get("variant"),d["gpu_launches"]))
except Exception as e: print("parse-fail",e)
PY
)" | tee -a "$OUT/summary.txt"; }
run c2 --steps 10 --warmup 3
run c4 --config c4 --steps 10 --warmup 3
run c3 --config c3 --steps 3 --warmup 3
run c5 --config c5 --per-gpu-share --steps 3 --warmup 3
run c2_v16 --steps 5 --warmup 3 --variant 16
run c2_v24 --steps 5 --warmup 3 --variant 24
