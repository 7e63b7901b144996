#!/bin/bash
# round-2 session 16 (8 GPUs): pipeline with pattern-aware shards at N = 8 / 4, larger job at N = 8, order 3 through Model
mkdir -p gpurun_out
L=gpurun_out/s16.log
: > $L
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 "$@"; }
for n in 8 4; do
  echo "== pipeline N=$n shared geometry, pattern-aware shards" >> $L
  tr $n bench.py --gpus $n --mode pipeline --pipeline-conforming 0 > gpurun_out/s16_pipeline_c0_n$n.json 2> gpurun_out/s16_pipeline_c0_n$n.err; echo "rc=$?" >> $L
  tail -1 gpurun_out/s16_pipeline_c0_n$n.json >> $L
done
echo "== pipeline N=8 shared geometry, 4000 depths (a job 4x as long)" >> $L
tr 8 bench.py --gpus 8 --mode pipeline --pipeline-conforming 0 --pipeline-depths 4000 > gpurun_out/s16_pipeline_c0_n8_4000.json 2> gpurun_out/s16_pipeline_c0_n8_4000.err; echo "rc=$?" >> $L
tail -1 gpurun_out/s16_pipeline_c0_n8_4000.json >> $L
cat $L
