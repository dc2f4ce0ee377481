#!/bin/bash
# Round-2 validation + measurement on one B200 (run under gpurun from the repository root): smoke, the GPU suite, both
# bench arms, the ncu launch list of one bench step and ncu --set full of the dominant kernels (K1, Cholesky, Schur
# pairs, block build, the KA kernel), the Cholesky time lines of both designs.  scripts/summarize_ncu.py turns the
# outputs into profiles/.
tag=${1:-r02_final}
out=gpurun_out; mkdir -p $out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee $out/pytest_gpu_$tag.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<PY
import json
d=json.load(open('$out/bench_$tag.json'))
print('value %.1fM ms %.3f steady %.3f'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step']))
print('e2e', json.dumps(d['e2e'])[:900])
print('surface', json.dumps(d.get('e2e_reference_surface'))[:700])
print({k:round(v['ms_per_step'],3) for k,v in d['stage_ms'].items()})
PY
tail -2 $out/bench_$tag.err
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_${tag}_reference.json 2> $out/bench_${tag}_reference.err; tail -c 400 $out/bench_${tag}_reference.json; tail -2 $out/bench_${tag}_reference.err
B="python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample-points 0"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_bench_$tag.log 2>&1
echo "== ncu --set full: K1, Cholesky, build/Schur, KA"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fm_eval_kernel -s 3 -c 2 -f -o $out/prof_k1_$tag $B > $out/ncu_full_k1_$tag.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:chol_persistent -s 1 -c 1 -f -o $out/prof_chol_$tag $B > $out/ncu_full_chol_$tag.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:ba_schur_pairs|ba_build_staged|ba_build_cam|ba_schur_prep|ba_point_inverse" -s 5 -c 5 -f -o $out/prof_schur_$tag $B > $out/ncu_full_schur_$tag.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:ka_solve_kernel -s 1 -c 1 -f -o $out/prof_ka_$tag python scripts/ka_throughput.py 4000 > $out/ncu_full_ka_$tag.log 2>&1
echo "== KA throughput"; timeout 300 python scripts/ka_throughput.py 4000 2>/dev/null | tail -1 | tee $out/ka_throughput_$tag.json
echo "== Cholesky time lines"
PXR_CHOL_TRACE=$out/chol_trace_$tag.txt timeout 300 $B > /dev/null 2>&1
PXR_CHOL_BAND=1 PXR_CHOL_TRACE=$out/chol_trace_band_$tag.txt timeout 300 $B > $out/bench_band_$tag.json 2>/dev/null; python - <<PY
import json
d=json.load(open('$out/bench_band_$tag.json')); print('band kernel: reduced solve', d['stage_ms']['reduced solve'])
PY
ls -la $out | tail -22
