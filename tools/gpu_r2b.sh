#!/bin/bash
# round 2, second GPU pass (2 GPUs): postnet + wire + peer-sink tests, bench at N=1 / N=2 (peer stores vs packed NCCL gather)
set -u
out=gpurun_out/r2b
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_postnet.py tests/test_gpu_wire.py -m gpu -q > $out/tests.txt 2>&1; echo "tests rc=$?" >> $out/tests.txt
tail -25 $out/tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err; echo "n1 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_n2_peer.json 2> $out/bench_n2_peer.err; echo "n2 peer rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-peer > $out/bench_n2_packed.json 2> $out/bench_n2_packed.err; echo "n2 packed rc=$?"
tail -3 $out/bench_n2_peer.err $out/bench_n2_packed.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d.get('ms_per_pass'), {k:(round(v['ms'],4) if v.get('ms') else None) for k,v in d.get('kernels',{}).items()}, d.get('gather_verified'), d['config']['parallelism'][:90], d.get('e2e',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
