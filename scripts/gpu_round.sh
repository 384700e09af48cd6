#!/bin/bash
# One gpurun call: smoke, GPU parity tests, per-variant bench lines, ncu launch list + full capture.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh <tag> [stages...]
set -u
TAG=${1:-r01}; shift || true
STAGES=${*:-"smoke tests bench ncu"}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.csv" 2>&1
for st in $STAGES; do
case $st in
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
  tail -5 "$OUT/smoke.log";;
tests)
  timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
  tail -15 "$OUT/pytest_gpu.log";;
bench)
  for v in ${VARIANTS:-3 4 5 6 7}; do for b in 128; do
    timeout 600 python bench.py --steps 5 --warmup 3 --variant $v --block $b --no-cpu --no-e2e > "$OUT/bench_v${v}_b${b}.json" 2> "$OUT/bench_v${v}_b${b}.err"
    echo "bench v$v b$b rc=$? $(python - "$OUT/bench_v${v}_b${b}.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print("value=%.3e ms/step=%.3f frac=%.3f seg/track=%.2f"%(d["value"],d["ms_per_step"],r["frac"],d["config"]["segments_per_track"]))
except Exception as e: print("parse-fail",e)
PY
)" | tee -a "$OUT/summary.txt"
  done; done
  timeout 900 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "bench default rc=$?" | tee -a "$OUT/summary.txt"
  cat "$OUT/bench_default.json"
  timeout 600 python bench.py --impl reference > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "bench reference rc=$?" | tee -a "$OUT/summary.txt"
  cat "$OUT/bench_reference.json";;
ncu)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches.csv" \
      python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e ${NCU_ARGS:-} > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?" | tee -a "$OUT/summary.txt"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:walk_ -s 2 -c 2 -o "$OUT/walk_full" -f \
      python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e ${NCU_ARGS:-} > "$OUT/ncu_full.log" 2>&1; echo "ncu full rc=$?" | tee -a "$OUT/summary.txt";;
esac
done
cat "$OUT/summary.txt"
