#!/bin/bash
# round-2 session 12: C1 (Example_01) bench mode, launch list + ncu full capture of the round-end configuration
mkdir -p gpurun_out
L=gpurun_out/s12.log
: > $L
echo "== example01 mode" >> $L
timeout 600 python bench.py --mode example01 > gpurun_out/s12_example01.json 2> gpurun_out/s12_example01.err; echo "rc=$?" >> $L; tail -2 gpurun_out/s12_example01.err >> $L; cat gpurun_out/s12_example01.json >> $L
echo "== smoke" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
echo "== ncu" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02b_launches_bench_5M.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-companions --contexts 1 > gpurun_out/s12_ncu_launch.log 2>&1; echo "launch list rc=$?" >> $L
python tools/summarize_launches.py gpurun_out/r02b_launches_bench_5M.csv gpurun_out/r02b_launches_bench_5M_summary.csv >> $L 2>&1
gzip -f gpurun_out/r02b_launches_bench_5M.csv
head -14 gpurun_out/r02b_launches_bench_5M_summary.csv >> $L
cat $L
