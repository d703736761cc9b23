#!/bin/bash
# parity, the clock trace and stage times of the rebuilt match+assemble kernel, the identity post-network kernel A/B
set -u
out=gpurun_out/${1:-r2j}
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt; tail -5 $out/tests.txt
timeout 300 python tools/trace_match_assemble.py 30 $out/trace_p30.json > $out/trace_p30.txt 2>&1; head -4 $out/trace_p30.txt
timeout 300 python tools/tune_r2.py 30 quick > $out/tune_p30.txt 2>&1; cat $out/tune_p30.txt | tail -3
for cfg in p30 net128; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 > $out/bench_$cfg.json 2> $out/bench_err_$cfg.txt; python - <<PY
import json
d=json.loads(open("$out/bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", d["value"], d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
PY
done
SPG_POST_IDENT=0 timeout 300 python bench.py --config net128 --steps 10 --warmup 3 > $out/bench_net128_generic.json 2>> $out/bench_err_net128.txt; python - <<PY
import json
d=json.loads(open("$out/bench_net128_generic.json").read().strip().splitlines()[-1])
print("net128 generic", d["value"], d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
PY
