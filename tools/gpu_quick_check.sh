#!/bin/bash
set -u
out=gpurun_out/${1:-r2q}
mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/tests_nms.txt 2>&1; echo "rc=$?" >> $out/tests_nms.txt; tail -3 $out/tests_nms.txt
if ! grep -q "rc=0" $out/tests_nms.txt; then echo "parity failed or hung: stop"; exit 1; fi
timeout 200 python tools/tune_r2.py 30 quick > $out/tune_p30.txt 2>&1; cut -c1-300 $out/tune_p30.txt | head -2
timeout 200 python tools/tune_r2.py 10 quick > $out/tune_p10.txt 2>&1; cut -c1-300 $out/tune_p10.txt | head -1
for cfg in p30 512; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 > $out/bench_$cfg.json 2> $out/bench_err_$cfg.txt; python - <<PY
import json
d=json.loads(open("$out/bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", d["value"], d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
PY
done
