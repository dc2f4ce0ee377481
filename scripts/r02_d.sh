#!/bin/bash
tag=${1:-r02d}
out=gpurun_out; mkdir -p $out
echo "== cholesky tests (band kernel)"; timeout 600 python -m pytest tests/test_gpu_ba_parity.py tests/test_gpu_edge_cases.py -q -m gpu -k "cholesky or bail_out" --durations=3 2>&1 | tail -8
echo "== Cholesky time line"
PXR_CHOL_TRACE=$out/chol_trace_$tag.txt timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample-points 0 > /dev/null 2>&1; awk 'NR%6==1' $out/chol_trace_$tag.txt | head -12; tail -2 $out/chol_trace_$tag.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<PY
import json
d=json.load(open('$out/bench_$tag.json'))
print('value %.1fM ms %.3f steady %.3f'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step']))
print('e2e', json.dumps(d['e2e'])[:1200])
print('surface', json.dumps(d.get('e2e_reference_surface'))[:1200])
print({k:round(v['ms_per_step'],3) for k,v in d['stage_ms'].items()})
print('cpu', d['cpu_baseline'] and d['cpu_baseline'].get('value'))
PY
tail -3 $out/bench_$tag.err
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee $out/pytest_gpu_$tag.log
