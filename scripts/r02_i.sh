#!/bin/bash
# Final validation: fused tensor-core Schur pair walk (default) vs the walk on a precomputed T; ordered by importance
# (the GPU budget of the round ends somewhere in here).
tag=${1:-r02i}
out=gpurun_out; mkdir -p $out
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest -m gpu -x (as the driver runs it)"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $out/pytest_gpu_$tag.log
ab() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-surface --cpu-sample-points 0 > $out/bench_${tag}_$name.json 2> $out/bench_${tag}_$name.err
  python - <<PY
import json
d=json.load(open('$out/bench_${tag}_$name.json'))
print('$name', 'ms %.3f steady %.3f'%(d['ms_per_step'],d['steady_state']['ms_per_step']), {k:round(v['ms_per_step'],4) for k,v in d['stage_ms'].items()})
PY
}
echo "== A/B"; ab fused PXR_X=1; ab mma PXR_SCHUR_KERNEL=mma
echo "== full bench"; timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<PY
import json
d=json.load(open('$out/bench_$tag.json'))
print('value %.1fM ms %.3f steady %.3f'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step']))
print('e2e', json.dumps({k:d['e2e'][k] for k in ('value','seconds','library_seconds','lm_loop_seconds','first_call_seconds','observations_refetched')}), 'full', d['e2e']['full_upload']['seconds'])
print('surface', json.dumps(d.get('e2e_reference_surface'))[:300])
PY
tail -2 $out/bench_$tag.err
echo "== ncu of the pair kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:schur_pairs -c 2 -f -o $out/schur_pairs_$tag python bench.py --steps 2 --warmup 1 --no-e2e --no-surface --cpu-sample-points 0 > $out/ncu_$tag.log 2>&1
ncu -i $out/schur_pairs_$tag.ncu-rep --page raw --csv > $out/schur_pairs_${tag}_raw.csv 2>/dev/null; rm -f $out/schur_pairs_$tag.ncu-rep
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-surface --cpu-sample-points 0 > /dev/null 2>&1
echo "== configs4 shard"; timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --workload configs4 --no-e2e --no-surface --cpu-sample-points 0 > $out/bench_${tag}_c4.json 2> $out/bench_${tag}_c4.err; tail -c 600 $out/bench_${tag}_c4.json
