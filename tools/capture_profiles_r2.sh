#!/bin/bash
# Run on the GPU box (gpurun): produces everything profiles/r2_final/ is built from.  One GPU.
set -u
out=gpurun_out/r2_final
mkdir -p $out
timeout 300 python bench.py --steps 20 --warmup 5 > $out/bench_p30.json 2> $out/bench_p30.err
for c in p10 f64 net128 512 imhn; do
  python bench.py --config $c --steps 20 --warmup 5 > $out/bench_$c.json 2> $out/bench_$c.err
done
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_512.csv \
    python bench.py --config 512 --steps 2 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_launches_512.log 2>&1
for k in nms_peaks limb_score match_assemble; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $out/prof_$k \
      python bench.py --steps 1 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:postnet -s 4 -c 1 -f -o $out/prof_postnet \
    python bench.py --config net128 --steps 1 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_postnet.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:limb_score -s 4 -c 1 -f -o $out/prof_limb_score_512 \
    python bench.py --config 512 --steps 1 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_limb_score_512.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:postnet -s 4 -c 1 -f -o $out/prof_postnet_512 \
    python bench.py --config 512 --steps 1 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_postnet_512.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nms_peaks -s 4 -c 1 -f -o $out/prof_nms_peaks_512 \
    python bench.py --config 512 --steps 1 --warmup 3 --passes 1 --no-cpu-baseline > $out/ncu_nms_peaks_512.log 2>&1
python tools/trace_match_assemble.py 30 > $out/trace_match_assemble.txt 2>&1
python tools/tune_r2.py 30 quick > $out/tune_p30.txt 2>&1
python tools/dropin_latency.py 128 30 > $out/dropin_latency_128.json 2>> $out/dropin.err
python tools/dropin_latency.py 512 14 > $out/dropin_latency_512.json 2>> $out/dropin.err
for t in memcheck synccheck; do
  compute-sanitizer --tool $t python tools/sanitize_run.py > $out/sanitizer_$t.txt 2>&1
done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.txt 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/gpu.csv
lscpu | grep -E "Model name|^CPU\(s\)" > $out/cpu.txt
ls -la $out
