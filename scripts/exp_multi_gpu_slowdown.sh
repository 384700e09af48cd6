#!/bin/bash
# Why is a c2 move 2.70 ms under torchrun on a multi-GPU box and 2.57 ms on a 1-GPU box?  Run on a 2-GPU box.
out=gpurun_out/r02w; mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.max.sm,clocks.max.mem,power.limit,ecc.mode.current --format=csv > $out/smi.txt
for g in 0 1; do
  CUDA_VISIBLE_DEVICES=$g timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $out/solo_gpu$g.json 2> $out/solo_gpu$g.err
done
# both at once, independent processes, no NCCL
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $out/pair_gpu0.json 2> $out/pair_gpu0.err &
p0=$!
CUDA_VISIBLE_DEVICES=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $out/pair_gpu1.json 2> $out/pair_gpu1.err &
p1=$!
wait $p0 $p1
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-extra > $out/torchrun_n2.json 2> $out/torchrun_n2.err
for f in $out/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('  ms/step %.3f'%j['ms_per_step'], j['roofline']['kernel'], j.get('clocks'))
except Exception as e: print('  failed', e)
PY
done
