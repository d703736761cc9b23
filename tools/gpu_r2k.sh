#!/bin/bash
# parity of the touched kernels, trace + stage times, ncu captures of the identity post-network kernel and match_assemble
set -u
out=gpurun_out/${1:-r2k}
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt; tail -3 $out/tests.txt
timeout 300 python tools/trace_match_assemble.py 30 > $out/trace_p30.txt 2>&1; head -3 $out/trace_p30.txt
timeout 300 python tools/tune_r2.py 30 quick > $out/tune_p30.txt 2>&1; cat $out/tune_p30.txt | tail -3
timeout 300 python bench.py --config p30 --steps 10 --warmup 3 > $out/bench_p30.json 2> $out/bench_err_p30.txt; python - <<PY
import json
d=json.loads(open("$out/bench_p30.json").read().strip().splitlines()[-1])
print("p30", d["value"], d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:postnet -s 4 -c 1 -f -o $out/prof_postnet_ident \
    python bench.py --config net128 --steps 1 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_postnet.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_assemble -s 4 -c 1 -f -o $out/prof_match_assemble \
    python bench.py --steps 1 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_ma.log 2>&1
ls -la $out
