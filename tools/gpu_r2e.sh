#!/bin/bash
# round 2, profile pass (1 GPU): full parity suite, drop-in latency, bench of every configuration, ncu captures
set -u
out=gpurun_out/r2_final
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt
tail -8 $out/tests.txt
timeout 200 python tools/dropin_latency.py 512 14 > $out/dropin_latency_512.json 2> $out/dropin_latency.err; cat $out/dropin_latency_512.json
timeout 200 python tools/dropin_latency.py 128 30 > $out/dropin_latency_128.json 2>> $out/dropin_latency.err; cat $out/dropin_latency_128.json
bash tools/capture_profiles_r2.sh > $out/capture.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_final/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d.get('ms_per_pass'), {k:(round(v['ms'],4) if v.get('ms') else None) for k,v in d.get('kernels',{}).items()}, 'e2e', round(d.get('e2e',{}).get('value',0)), 'grouping_only', d.get('grouping_only',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
ls -la $out | head -50
