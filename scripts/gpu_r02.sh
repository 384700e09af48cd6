#!/bin/bash
# Round-2 GPU call helper.  Usage: bash scripts/gpu_r02.sh <tag> <stage> [...]
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for st in "$@"; do
case $st in
smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/smoke.log";;
tests) timeout 2400 python -m pytest tests -m gpu -x -q --timeout=900 ${PYTEST_ARGS:-} > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -15 "$OUT/pytest_gpu.log";;
e2e) timeout 900 python scripts/exp_e2e.py ${E2E_MODES:-} > "$OUT/e2e.log" 2>&1; echo "e2e rc=$?" | tee -a "$OUT/summary.txt"; cat "$OUT/e2e.log";;
bench) timeout 1200 python bench.py ${BENCH_ARGS:-} > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"; cat "$OUT/bench.json"; tail -3 "$OUT/bench.err";;
benchref) timeout 900 python bench.py --impl reference > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "benchref rc=$?" | tee -a "$OUT/summary.txt"; cat "$OUT/bench_reference.json";;
esac
done
