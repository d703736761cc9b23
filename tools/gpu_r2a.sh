#!/bin/bash
# round 2, first GPU pass: parity tests, the four bench configurations, the reference arm.  One GPU.
set -u
out=gpurun_out/r2a
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu.csv
timeout 900 python -m pytest tests -m gpu -x -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt
tail -5 $out/tests.txt
for c in p30 p10 f64 512; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 5 > $out/bench_$c.json 2> $out/bench_$c.err; echo "bench $c rc=$?"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d.get('ms_per_pass'), {k:round(v['ms'],4) for k,v in d.get('kernels',{}).items()}, d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
