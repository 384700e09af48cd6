#!/bin/bash
# GPU experiment: compact layout + edge-function walk (variants 20/21/22) vs the plane-record defaults.
set -u
OUT=gpurun_out/${1:-edge}
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -x -k "20 or 21 or 22 or 23 or edge_walk" > "$OUT/pytest_edge.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_edge.log"
run() {  # name, args...
  local name=$1; shift
  timeout 900 python bench.py --no-cpu --no-e2e "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  echo "$name rc=$? $(python - "$OUT/bench_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print("value=%.3e ms/step=%.3f frac=%.3f variant=%s"%(d["value"],d["ms_per_step"],r["frac"],d["config"].get("variant")))
except Exception as e: print("parse-fail",e)
PY
)" | tee -a "$OUT/summary.txt"
}
for v in 8 20 22 23 21; do run c2_v$v --steps 5 --warmup 3 --variant $v; done
for v in 8 20; do run c3_v$v --config c3 --steps 3 --warmup 3 --variant $v; done
for v in 16 21; do run c5_v$v --config c5 --per-gpu-share --steps 3 --warmup 3 --variant $v; done
