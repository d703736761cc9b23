#!/bin/bash
# two GPUs: the whole GPU suite (the NVLink gather tests run), then the 2-GPU bench line
set -u
out=gpurun_out/${1:-r2n}
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt; tail -3 $out/tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_p30_n2.json 2> $out/bench_n2_err.txt; tail -c 600 $out/bench_p30_n2.json
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 > $out/bench_p30_n1.json 2> $out/bench_n1_err.txt; python - <<PY
import json
for n in (1, 2):
    d=json.loads(open("$out/bench_p30_n%d.json" % n).read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], d.get("gather"), {k:round(v["ms"],4) for k,v in d.get("kernels",{}).items()})
PY
