#!/bin/bash
set -u
out=gpurun_out/tune
mkdir -p $out
timeout 600 python tools/tune_r2.py 30 > $out/tune_p30.txt 2> $out/tune_p30.err; cat $out/tune_p30.txt; tail -3 $out/tune_p30.err
