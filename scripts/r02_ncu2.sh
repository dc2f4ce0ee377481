#!/bin/bash
# ncu --set full of the dominant kernels; every report is reduced to its raw-page CSV on the box (gpurun_out/ is capped at
# 64 MiB), only K1's .ncu-rep (source page) is kept.
tag=${1:-r02}
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample-points 0"
cap() { name=$1; kern=$2; skip=$3; cnt=$4; shift 4
  timeout 500 ncu --set full --clock-control none --import-source on -k "$kern" -s $skip -c $cnt -f -o /tmp/prof_$name "$@" > $out/ncu_full_${name}_$tag.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $out/prof_${name}_$tag.csv 2>/dev/null; tail -2 $out/ncu_full_${name}_$tag.log; }
cap k1 regex:fm_eval_kernel 3 2 $B; cp /tmp/prof_k1.ncu-rep $out/prof_k1_$tag.ncu-rep
ncu -i /tmp/prof_k1.ncu-rep --page source --csv 2>/dev/null | head -3000 > $out/prof_k1_source_$tag.csv
cap chol regex:chol_persistent 1 1 $B
cap schur "regex:ba_schur_pairs|ba_build_staged|ba_build_cam|ba_schur_prep|ba_point_inverse" 5 5 $B
cap ka regex:ka_solve_kernel 1 1 python scripts/ka_throughput.py 4000
PXR_CHOL_BAND=1 cap band regex:chol_band 1 1 $B
du -sh $out; ls -la $out | tail -14
