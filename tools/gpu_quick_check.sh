#!/bin/bash
# quick check after a kernel change: the whole GPU suite, stage times alone (f32 and f32-evaluated-in-f64), two bench lines
set -u
out=gpurun_out/${1:-quick}
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests.txt 2>&1; echo "rc=$?" >> $out/tests.txt; tail -3 $out/tests.txt
if ! grep -q "rc=0" $out/tests.txt; then echo "GPU suite failed or hung: stop"; exit 1; fi
timeout 200 python tools/tune_r2.py 30 quick > $out/tune_p30.txt 2>&1; cut -c1-300 $out/tune_p30.txt | head -1
for cfg in p30 net128 f64; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 > $out/bench_$cfg.json 2> $out/bench_err_$cfg.txt; python - <<PY
import json
d=json.loads(open("$out/bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", d["value"], d["ms_per_pass"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
PY
done
