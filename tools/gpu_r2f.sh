#!/bin/bash
# round 2, quick pass (1 GPU): post-network kernel tests + pipeline configs + p30
set -u
out=gpurun_out/r2f
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_postnet.py tests/test_gpu_dropin.py tests/test_gpu_wire.py -m gpu -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt
tail -8 $out/tests.txt
for c in net128 512 imhn p30; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_$c.json 2> $out/bench_$c.err; echo "bench $c rc=$?"; tail -2 $out/bench_$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d.get('ms_per_pass'), {k:(round(v['ms'],4) if v.get('ms') else None) for k,v in d.get('kernels',{}).items()}, 'e2e', round(d.get('e2e',{}).get('value',0)), 'grouping_only', d.get('grouping_only',{}).get('value'), 'frac', {k:(round(v['frac_of_hbm_peak'],3) if v.get('frac_of_hbm_peak') else None) for k,v in d.get('kernels',{}).items()})
    except Exception as e: print(f,'ERR',e)
PY
