#!/bin/bash
# last single-GPU check of the round: memcheck of the small run, score-filter cost, full GPU suite, default bench + reference arm
out=gpurun_out/r02x; mkdir -p $out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > $out/memcheck.txt 2>&1; echo "memcheck rc=$?"
tail -3 $out/memcheck.txt
timeout 300 python scripts/exp_score_bins.py > $out/score_bins.jsonl 2> $out/score_bins.err; echo "score_bins rc=$?"; cat $out/score_bins.jsonl
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee $out/pytest_gpu_tail.txt
timeout 400 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference rc=$?"
timeout 600 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
python -c "from __graft_entry__ import smoke; smoke(); print('smoke ok')" 2>&1 | tail -2
