#!/bin/bash
# round-2 session 24 (8 GPUs): the step bench at N = 1 / 2 / 4 / 8 as the driver launches it (weak scaling)
mkdir -p gpurun_out
L=gpurun_out/s24.log
: > $L
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/s24_n1.json 2> gpurun_out/s24_n1.err; echo "N=1 rc=$?" >> $L
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s24_n$n.json 2> gpurun_out/s24_n$n.err; echo "N=$n rc=$?" >> $L
done
python - >> $L <<PY
import json
base = None
for n in (1, 2, 4, 8):
    d = json.load(open('gpurun_out/s24_n%d.json' % n))
    base = base or d['value']
    print('N=%d value %.2f e2e %.2f ms/step %.1f efficiency %.3f' % (n, d['value'], d['e2e']['value'], d['ms_per_step'], d['value'] / (n * base)))
PY
cat $L
