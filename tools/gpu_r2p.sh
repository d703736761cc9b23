#!/bin/bash
# banded persistent nms: parity first (short timeout), stage times, then the screen-sample sweep of limb_score
set -u
out=gpurun_out/${1:-r2p}
mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/tests_nms.txt 2>&1; echo "rc=$?" >> $out/tests_nms.txt; tail -3 $out/tests_nms.txt
if ! grep -q "rc=0" $out/tests_nms.txt; then echo "nms parity failed or hung: stop"; exit 1; fi
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt; tail -3 $out/tests.txt
for v in "" _s6 _s8 _s12; do
  SPG_LIB=$PWD/improved_body_parts_b200/libspgroup$v.so timeout 200 python tools/tune_r2.py 30 quick 2>&1 | head -1 >> $out/tune_variants.txt
done
cat $out/tune_variants.txt
for cfg in p30 512; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 > $out/bench_$cfg.json 2> $out/bench_err_$cfg.txt; python - <<PY
import json
d=json.loads(open("$out/bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", d["value"], d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
PY
done
