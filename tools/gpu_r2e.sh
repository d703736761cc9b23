#!/bin/bash
# round 2, profile pass (1 GPU): full parity suite, drop-in latency, bench of every configuration, ncu captures
set -u
out=gpurun_out/r2_final
mkdir -p $out
timeout 300 python __graft_entry__.py smoke > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.txt
timeout 1200 python -m pytest tests -m gpu -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt
tail -8 $out/tests.txt
timeout 200 python tools/dropin_latency.py 512 14 > $out/dropin_latency_512.json 2> $out/dropin_latency.err; cat $out/dropin_latency_512.json
timeout 200 python tools/dropin_latency.py 128 30 > $out/dropin_latency_128.json 2>> $out/dropin_latency.err; cat $out/dropin_latency_128.json
bash tools/capture_profiles_r2.sh > $out/capture.log 2>&1
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_run.py > $out/sanitizer_$tool.txt 2>&1; echo "sanitizer $tool rc=$?"; tail -3 $out/sanitizer_$tool.txt
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_final/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d.get('ms_per_pass'), {k:(round(v['ms'],4) if v.get('ms') else None) for k,v in d.get('kernels',{}).items()}, 'e2e', round(d.get('e2e',{}).get('value',0)), 'grouping_only', d.get('grouping_only',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
ls -la $out | head -50
