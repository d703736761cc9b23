#!/bin/bash
set -u
out=gpurun_out/r2_final
mkdir -p $out
python bench.py --config 512 --steps 20 --warmup 5 > $out/bench_512.json 2> $out/bench_512.err
ncu --set full --clock-control none --import-source on -k regex:postnet -s 4 -c 1 -f -o $out/prof_postnet_512 \
    python bench.py --config 512 --steps 1 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_postnet_512.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_512.csv \
    python bench.py --config 512 --steps 2 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_launches_512.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.txt 2>&1; tail -1 $out/smoke.txt
