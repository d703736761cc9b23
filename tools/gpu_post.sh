#!/bin/bash
set -u
out=gpurun_out/${1:-post}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_postnet.py tests/test_gpu_dropin.py -m gpu -x -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt; tail -3 $out/tests.txt
for cfg in 512; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 > $out/bench_$cfg.json 2> $out/bench_err_$cfg.txt; python - <<PY
import json
d=json.loads(open("$out/bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", d["value"], d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()}, d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline"]["traffic"])
PY
done
