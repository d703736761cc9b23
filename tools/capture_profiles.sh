#!/bin/bash
# Run on the GPU box (gpurun): produces everything profiles/r1_final/ is built from.  One GPU.
set -u
out=gpurun_out/final
mkdir -p $out
python bench.py --steps 50 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
for k in nms_peaks limb_score limb_match assemble; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/prof_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_$k.log 2>&1
done
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu.csv
lscpu | grep -E "Model name|^CPU\(s\)" > $out/cpu.txt
ls -la $out
