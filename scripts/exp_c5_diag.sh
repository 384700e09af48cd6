#!/bin/bash
# c5 per-GPU share: walk-kernel time / DRAM bytes / L2 hit rate for processing-order and ticket-size choices
OUT=gpurun_out/${1:-c5diag}; mkdir -p $OUT
for cfg in "morton=0 claim_run=4" "morton=1 claim_run=4" "morton=1 claim_run=1" "morton=0 claim_run=1"; do
  tag=$(echo $cfg | tr ' =' '__')
  opts=""; for o in $cfg; do opts="$opts --opt $o"; done
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,l1tex__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:walk_persist -s 3 -c 1 --csv --log-file $OUT/$tag.csv \
     python bench.py --config c5 --per-gpu-share --steps 2 --warmup 1 --no-cpu --no-e2e $opts > $OUT/$tag.log 2>&1
  echo "$cfg: $(python - $OUT/$tag.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; im=hdr.index('Metric Name'); iv=hdr.index('Metric Value')
print({r[im]:r[iv] for r in rows[1:]})
PY
)"
done
