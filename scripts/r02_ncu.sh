#!/bin/bash
# ncu captures of the final kernels with outputs small enough to travel back (gpurun_out/ is capped at 64 MiB): every
# report is turned into its raw-page CSV on the box and only K1's .ncu-rep is kept.
tag=${1:-r02}
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample-points 0"
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_deterministic.py tests/test_gpu_ba_parity.py tests/test_gpu_mirror.py tests/test_gpu_zz_baseline_configs.py tests/test_refine_hloc.py tests/test_gpu_block_mode.py -q -m gpu 2>&1 | tail -12 | tee $out/pytest_new_$tag.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<PY
import json
d=json.load(open('$out/bench_$tag.json'))
print('value %.1fM ms %.3f steady %.3f'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step']))
print('e2e', json.dumps(d['e2e'])[:1100])
print('surface', json.dumps(d.get('e2e_reference_surface'))[:500])
PY
tail -2 $out/bench_$tag.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_bench_$tag.log 2>&1
echo "== ncu --set full: K1, Cholesky, build/Schur, KA"
cap() { name=$1; shift; timeout 500 ncu --set full --clock-control none --import-source on "$@" -f -o /tmp/prof_$name > $out/ncu_full_${name}_$tag.log 2>&1; ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $out/prof_${name}_$tag.csv 2>/dev/null; }
cap k1 -k regex:fm_eval_kernel -s 3 -c 2 $B; cp /tmp/prof_k1.ncu-rep $out/prof_k1_$tag.ncu-rep
ncu -i /tmp/prof_k1.ncu-rep --page source --csv 2>/dev/null | head -4000 > $out/prof_k1_source_$tag.csv
cap chol -k regex:chol_persistent -s 1 -c 1 $B
cap schur -k "regex:ba_schur_pairs|ba_build_staged|ba_build_cam|ba_schur_prep|ba_point_inverse" -s 5 -c 5 $B
cap ka -k regex:ka_solve_kernel -s 1 -c 1 python scripts/ka_throughput.py 4000
echo "== Cholesky time lines"
PXR_CHOL_TRACE=$out/chol_trace_$tag.txt timeout 300 $B > /dev/null 2>&1
PXR_CHOL_BAND=1 PXR_CHOL_TRACE=$out/chol_trace_band_$tag.txt timeout 300 $B > $out/bench_band_$tag.json 2>/dev/null
du -sh $out; ls -la $out | tail -20
