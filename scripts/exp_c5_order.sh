#!/bin/bash
# c5 per-GPU share: processing order of the binned walk (z-major slabs vs Morton blocks) -- per-kernel time, DRAM bytes, L2 hit rate
OUT=gpurun_out/${1:-c5order}; mkdir -p $OUT
for m in 0 1; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum --clock-control none --csv --log-file $OUT/launches_morton$m.csv \
     python bench.py --config c5 --per-gpu-share --steps 2 --warmup 1 --no-cpu --no-e2e --opt morton=$m > $OUT/ncu_morton$m.log 2>&1
  echo "morton=$m rc=$?"
  python - $OUT/launches_morton$m.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; 
ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[iid],r[ik][:60]),{})[r[im]]=r[iv]
for (i,k),m in list(d.items())[-14:]:
    print(i,k,m)
PY
done
