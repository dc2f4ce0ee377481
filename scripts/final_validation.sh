#!/bin/bash
# Round-end validation + measurement on one B200 (run under gpurun from the repository root):
#   smoke, the GPU parity suite, both bench arms, an e2e A/B of the upload stream, the ncu launch list, ncu --set full of
#   the dominant kernels and the panel-CTA time line of the persistent Cholesky.  Outputs go to gpurun_out/ with the
#   given tag; scripts/summarize_ncu.py turns them into profiles/.
tag=${1:-r01_final}
out=gpurun_out
mkdir -p $out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 600 $out/bench_$tag.json
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference > $out/bench_${tag}_reference.json 2>> $out/bench_$tag.err; tail -c 300 $out/bench_${tag}_reference.json
echo "== e2e with the slab on the context stream (A/B)"
PXR_UPLOAD_SAME_STREAM=1 timeout 600 python bench.py --cpu-sample-points 0 > $out/bench_${tag}_samestream.json 2>> $out/bench_$tag.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample-points 0 > $out/ncu_bench_$tag.log 2>&1
B="python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample-points 0"
echo "== ncu --set full: K1, Cholesky, build/Schur"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fm_eval_kernel -s 3 -c 2 -f -o $out/prof_k1_$tag $B > $out/ncu_full_k1_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_persistent -s 1 -c 1 -f -o $out/prof_chol_$tag $B > $out/ncu_full_chol_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:ba_schur_pairs|ba_build_kernel|ba_build_cam|ba_schur_prep|ba_point_inverse" -s 5 -c 5 -f -o $out/prof_schur_$tag $B > $out/ncu_full_schur_$tag.log 2>&1
echo "== Cholesky time line"
PXR_CHOL_TRACE=$out/chol_trace_$tag.txt timeout 300 $B > /dev/null 2>&1
ls -la $out | tail -15
