#!/bin/bash
# parity first, then the clock trace and the stage times of the rebuilt match+assemble kernel
set -u
out=gpurun_out/${1:-r2i}
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt; tail -5 $out/tests.txt
timeout 300 python tools/trace_match_assemble.py 30 $out/trace_p30.json > $out/trace_p30.txt 2>&1; head -4 $out/trace_p30.txt
timeout 300 python tools/tune_r2.py 30 quick > $out/tune_p30.txt 2>&1; cat $out/tune_p30.txt | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 > $out/bench_p30.json 2> $out/bench_err.txt; python - <<PY
import json
d=json.loads(open("$out/bench_p30.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
PY
