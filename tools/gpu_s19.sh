#!/bin/bash
# round-2 session 19 (2 GPUs): run-to-run spread of the headline at the driver's step count, N = 2 under torchrun
mkdir -p gpurun_out
L=gpurun_out/s19.log
: > $L
for i in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/s19_run$i.json 2> gpurun_out/s19_run$i.err
  python - $i >> $L <<PY
import json, sys
d = json.load(open('gpurun_out/s19_run%s.json' % sys.argv[1]))
print('run', sys.argv[1], 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],3), 'clocks', d['clocks'])
PY
done
echo "== N=2" >> $L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s19_n2.json 2> gpurun_out/s19_n2.err; echo "rc=$?" >> $L
python - >> $L <<PY
import json
d = json.load(open('gpurun_out/s19_n2.json'))
print('N=2 value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],1))
PY
echo "== reference arm under torchrun N=2 (rank 0 only)" >> $L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s19_ref_n2.json 2> gpurun_out/s19_ref_n2.err; echo "rc=$?" >> $L; cut -c1-200 gpurun_out/s19_ref_n2.json >> $L
cat $L
