#!/bin/bash
# quick pass (1 GPU): parity suite + tuning sweep + p30 bench of the current build
set -u
out=gpurun_out/r2g
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wire.py -m gpu -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt
tail -4 $out/tests.txt
timeout 600 python tools/tune_r2.py 30 > $out/tune_p30.txt 2> $out/tune_p30.err; cat $out/tune_p30.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_p30.json 2> $out/bench_p30.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g/bench_p30.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_pass'], {k:round(v['ms'],4) for k,v in d['kernels'].items()}, d['roofline']['limb_score_frac'], d['roofline']['nms_peaks_frac'], d['roofline']['traffic'])
PY
