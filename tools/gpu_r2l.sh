#!/bin/bash
set -u
out=gpurun_out/${1:-r2l}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_postnet.py tests/test_gpu_dropin.py -m gpu -x -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt; tail -3 $out/tests.txt
for cfg in net128 imhn; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 > $out/bench_$cfg.json 2> $out/bench_err_$cfg.txt; python - <<PY
import json
d=json.loads(open("$out/bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", d["value"], d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:postnet -s 4 -c 1 -f -o $out/prof_postnet_ident \
    python bench.py --config net128 --steps 1 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_postnet.log 2>&1
