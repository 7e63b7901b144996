#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/final2.log
: > $L
for pf in 1 0; do
REMO_EBE_PREFETCH=$pf REMO_PROBE_SIZE=5M timeout 400 python tools/spmm_probe.py --ks 5,1 2>&1 | grep "^k=\|rror" | sed "s/^/pf=$pf /" >> $L
done
a=$(grep "pf=1 k=5" $L | awk '{print $4}'); b=$(grep "pf=0 k=5" $L | awk '{print $4}')
best=$(python -c "print(1 if float('$a') <= float('$b') else 0)")
echo "best prefetch=$best ($a vs $b)" >> $L
export REMO_EBE_PREFETCH=$best
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
python bench.py > gpurun_out/bench_default_ebe2.json 2> gpurun_out/bench_default_ebe2.err; echo "bench rc=$?" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_5M_ebe2.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --contexts 1 > gpurun_out/ncu_launch_ebe2.log 2>&1; echo "launch list rc=$?" >> $L
REMO_PROBE_SIZE=5M timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_spmm_ebe -s 30 -c 2 -o gpurun_out/ebe2_k5_5M -f \
  python tools/spmm_probe.py --ks 5 > gpurun_out/ncu_ebe2_5M.log 2>&1; echo "ncu full rc=$?" >> $L
python -c "
import json
d=json.load(open('gpurun_out/bench_default_ebe2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['iterations'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['kernel'][:20], d.get('cpu_baseline',{}).get('value'))" >> $L 2>&1
cat $L
