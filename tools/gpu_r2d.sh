#!/bin/bash
# round 2, multi-GPU pass (8 GPUs): 2-GPU gather tests, scaling 1/2/4/8 with NVLink peer stores, packed NCCL gather at 8 for comparison
set -u
out=gpurun_out/r2d
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_wire.py -m gpu -q -k "two_gpu" > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt
tail -5 $out/tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err; echo "n1 rc=$?"
port=29520
for n in 2 4 8; do
  port=$((port+1))
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 --trace $out/timeline_n$n.json > $out/bench_n$n.json 2> $out/bench_n$n.err; echo "n$n rc=$?"
done
port=$((port+1))
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 20 --warmup 5 --no-peer --trace $out/timeline_n8_packed.json > $out/bench_n8_packed.json 2> $out/bench_n8_packed.err; echo "n8 packed rc=$?"
for f in $out/bench_n*.err; do echo "== $f"; tail -n 3 $f; done
python - <<'PY'
import json,glob
base=None
for f in sorted(glob.glob('gpurun_out/r2d/bench_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if d['n_gpus']==1: base=d['value']
        print(f, d['n_gpus'], round(d['value']), round(d['ms_per_pass'],4), 'eff', round(d['value']/d['n_gpus']/base,3) if base else None, {k:(round(v['ms'],4) if v.get('ms') else None) for k,v in d.get('kernels',{}).items()}, d.get('gather_verified'), 'e2e', round(d['e2e']['value']))
    except Exception as e: print(f,'ERR',e)
PY
