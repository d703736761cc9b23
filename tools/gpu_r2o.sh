#!/bin/bash
# screen-sample sweep of limb_score (build variants), stage times alone
set -u
out=gpurun_out/${1:-r2o}
mkdir -p $out
for v in "" _s6 _s8 _s12; do
  SPG_LIB=$PWD/improved_body_parts_b200/libspgroup$v.so timeout 300 python tools/tune_r2.py 30 quick 2>&1 | head -1 >> $out/tune_variants.txt
done
cat $out/tune_variants.txt
