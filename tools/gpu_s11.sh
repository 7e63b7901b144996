#!/bin/bash
# round-2 session 11 (8 GPUs): Model.simulate_logs pipeline at N = 8 / 4 / 2 (shared geometry), N = 8 with per-task interfaces,
# and the step bench at N = 8
mkdir -p gpurun_out
L=gpurun_out/s11.log
: > $L
nvidia-smi -L | wc -l >> $L
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
for n in 8 4 2; do
  echo "== pipeline N=$n shared geometry" >> $L
  tr $n bench.py --gpus $n --mode pipeline --pipeline-conforming 0 > gpurun_out/s11_pipeline_c0_n$n.json 2> gpurun_out/s11_pipeline_c0_n$n.err; echo "rc=$?" >> $L
  tail -1 gpurun_out/s11_pipeline_c0_n$n.json >> $L
done
echo "== pipeline N=8 per-task interfaces" >> $L
tr 8 bench.py --gpus 8 --mode pipeline --pipeline-conforming 1 > gpurun_out/s11_pipeline_c1_n8.json 2> gpurun_out/s11_pipeline_c1_n8.err; echo "rc=$?" >> $L
tail -1 gpurun_out/s11_pipeline_c1_n8.json >> $L
echo "== pipeline N=8 shared geometry, 1M size class" >> $L
tr 8 bench.py --gpus 8 --mode pipeline --pipeline-conforming 0 --pipeline-size 1M --pipeline-depths 400 > gpurun_out/s11_pipeline_c0_n8_1M.json 2> gpurun_out/s11_pipeline_c0_n8_1M.err; echo "rc=$?" >> $L
tail -1 gpurun_out/s11_pipeline_c0_n8_1M.json >> $L
echo "== step bench N=8" >> $L
tr 8 bench.py --gpus 8 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/s11_bench_n8.json 2> gpurun_out/s11_bench_n8.err; echo "rc=$?" >> $L
tail -1 gpurun_out/s11_bench_n8.json | cut -c1-400 >> $L
cat $L
