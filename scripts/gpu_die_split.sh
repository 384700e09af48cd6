#!/bin/bash
out=gpurun_out/r02z; mkdir -p $out
timeout 200 python -m pytest tests -q -m gpu -x -k "c1_parity or long_axial or unstructured_parity or score_bins" 2>&1 | tail -3
export EXP_STEPS=6
timeout 120 python scripts/exp_variants.py c4 0 24::die_split=0 24::die_split=1 24::die_split=0 24::die_split=1 2>/dev/null | grep "^{" | tee $out/die_split_c4.jsonl
timeout 200 python scripts/exp_variants.py c5 12500000 16::die_split=0 16::die_split=1 16::die_split=0 16::die_split=1 2>/dev/null | grep "^{" | tee $out/die_split_c5.jsonl
timeout 200 python scripts/exp_variants.py c2 0 24::die_split=0 24::die_split=1 16::die_split=0 16::die_split=1 8 2>/dev/null | grep "^{" | tee $out/die_split_c2.jsonl
