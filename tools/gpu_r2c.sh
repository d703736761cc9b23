#!/bin/bash
# round 2, third GPU pass (1 GPU): full parity suite with the fused match+assemble kernel, tuning sweep, bench configs
set -u
out=gpurun_out/r2c
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt
tail -12 $out/tests.txt
timeout 700 python tools/tune_r2.py 30 > $out/tune_p30.txt 2> $out/tune_p30.err; echo "tune rc=$?"
cat $out/tune_p30.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_p30.json 2> $out/bench_p30.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --unfused > $out/bench_p30_unfused.json 2> $out/bench_p30_unfused.err; echo "bench unfused rc=$?"
for c in f64 net128 512; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 5 > $out/bench_$c.json 2> $out/bench_$c.err; echo "bench $c rc=$?"; tail -2 $out/bench_$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d.get('ms_per_pass'), {k:(round(v['ms'],4) if v.get('ms') else None) for k,v in d.get('kernels',{}).items()}, 'e2e', round(d.get('e2e',{}).get('value',0)), 'grouping_only', d.get('grouping_only',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
