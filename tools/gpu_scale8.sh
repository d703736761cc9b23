#!/bin/bash
# eight GPUs of one box: the N = 8 bench line and the N = 1 line of the same box (weak scaling, NVLink peer-store gather)
set -u
out=gpurun_out/${1:-scale8}
mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > $out/bench_p30_n8.json 2> $out/bench_n8_err.txt
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 > $out/bench_p30_n1.json 2> $out/bench_n1_err.txt
python - <<PY
import json
for n in (1, 8):
    d=json.loads(open("$out/bench_p30_n%d.json" % n).read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_pass"], d.get("gather_verified"), {k:round(v["ms"],4) for k,v in d.get("kernels",{}).items()})
PY
