#!/bin/bash
# Validation + A/B of the tensor-core Schur pair walk (default) against the staged and the direct kernels.
tag=${1:-r02h}
out=gpurun_out; mkdir -p $out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest -m gpu -x (as the driver runs it)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $out/pytest_gpu_$tag.log
ab() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-surface --cpu-sample-points 0 > $out/bench_${tag}_$name.json 2> $out/bench_${tag}_$name.err
  python - <<PY
import json
d=json.load(open('$out/bench_${tag}_$name.json'))
print('$name', 'ms %.3f steady %.3f'%(d['ms_per_step'],d['steady_state']['ms_per_step']), {k:round(v['ms_per_step'],4) for k,v in d['stage_ms'].items()})
PY
}
echo "== A/B"; ab mma3 PXR_X=1; ab mma4 PXR_SCHUR_CTAS=4; ab staged PXR_SCHUR_KERNEL=staged
ab4() {  # configs[4] shard on one GPU: block-sparse path (sp_schur_pairs_kernel)
  name=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --workload configs4 --no-e2e --no-surface --cpu-sample-points 0 > $out/bench_${tag}_c4_$name.json 2> $out/bench_${tag}_c4_$name.err
  python - <<PY
import json
d=json.load(open('$out/bench_${tag}_c4_$name.json'))
print('configs4 shard $name', 'ms %.3f steady %.3f'%(d['ms_per_step'],d['steady_state']['ms_per_step']), {k:round(v['ms_per_step'],4) for k,v in d['stage_ms'].items()})
PY
}
echo "== A/B configs4 shard"; ab4 mma PXR_X=1; ab4 staged PXR_SCHUR_KERNEL=staged
echo "== parity subset with the staged kernel"; PXR_SCHUR_KERNEL=staged timeout 600 python -m pytest tests/test_gpu_ba_parity.py tests/test_gpu_deterministic.py -x -q 2>&1 | tail -3
echo "== ncu of the pair kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:schur_pairs -c 2 -f -o $out/schur_pairs_$tag python bench.py --steps 2 --warmup 1 --no-e2e --no-surface --cpu-sample-points 0 > $out/ncu_$tag.log 2>&1
ncu -i $out/schur_pairs_$tag.ncu-rep --page raw --csv > $out/schur_pairs_${tag}_raw.csv 2>/dev/null; rm -f $out/schur_pairs_$tag.ncu-rep
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-surface --cpu-sample-points 0 > /dev/null 2>&1
echo "== configs3 shape"; timeout 300 python scripts/bench_configs3.py 40000 2>$out/bench_configs3_$tag.err | tail -1 | tee $out/bench_configs3_$tag.json | cut -c1-900
echo "== full bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<PY
import json
d=json.load(open('$out/bench_$tag.json'))
print('value %.1fM ms %.3f steady %.3f'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step']))
print('e2e', json.dumps({k:d['e2e'][k] for k in ('value','seconds','library_seconds','lm_loop_seconds','first_call_seconds','observations_refetched')}), 'full', d['e2e']['full_upload']['seconds'])
print('surface', json.dumps(d.get('e2e_reference_surface'))[:300])
PY
tail -2 $out/bench_$tag.err
