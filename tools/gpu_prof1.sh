#!/bin/bash
set -u
out=gpurun_out/prof1
mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:postnet -s 4 -c 1 -f -o $out/prof_postnet \
    python bench.py --config net128 --steps 1 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_postnet.log 2>&1
ls -la $out
