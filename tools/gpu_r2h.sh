#!/bin/bash
set -u
out=gpurun_out/r2h
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt; tail -3 $out/tests.txt
timeout 200 python tools/dropin_latency.py 512 14 > $out/dropin_latency_512.json 2> $out/err.txt; cat $out/dropin_latency_512.json
timeout 200 python tools/dropin_latency.py 128 30 > $out/dropin_latency_128.json 2>> $out/err.txt; cat $out/dropin_latency_128.json
tail -3 $out/err.txt
