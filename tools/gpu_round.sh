#!/bin/bash
# One GPU round of evidence: parity tests, the default bench line, the ncu launch list of one bench step and a full
# capture of the PCG product kernel.  Run as:  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
mkdir -p gpurun_out
L=gpurun_out/round.log
: > $L
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" >> $L
REMO_BENCH_MESH_IMPROVE=0 python bench.py --no-cpu-baseline > gpurun_out/bench_default_plain_mesh.json 2> gpurun_out/bench_default_plain_mesh.err; echo "bench (plain mesh) rc=$?" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_5M.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --contexts 1 > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?" >> $L
REMO_PROBE_SIZE=5M timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_spmm_ebe -s 30 -c 2 -o gpurun_out/ebe_k5_5M -f \
  python tools/spmm_probe.py --ks 5 > gpurun_out/ncu_ebe_5M.log 2>&1; echo "ncu full rc=$?" >> $L
python -c "
import json
for f in ('bench_default', 'bench_default_plain_mesh'):
    d = json.load(open('gpurun_out/%s.json' % f))
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['iterations'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])" >> $L 2>&1
cat $L
