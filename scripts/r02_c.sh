#!/bin/bash
tag=${1:-r02c}
out=gpurun_out; mkdir -p $out
echo "== cholesky tests (band kernel)"; timeout 600 python -m pytest tests/test_gpu_ba_parity.py tests/test_gpu_edge_cases.py -q -m gpu -k "cholesky or full_solve or bail_out" --durations=5 2>&1 | tail -15
echo "== residency + upload tests"; timeout 600 python -m pytest tests/test_gpu_resident.py tests/test_gpu_upload.py -q -m gpu 2>&1 | tail -15
echo "== bench (band Cholesky, window e2e)"; timeout 700 python bench.py --steps 20 --warmup 5 --cpu-sample-points 0 > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<PY
import json
d=json.load(open('$out/bench_$tag.json'))
print('value %.1fM ms %.3f steady %.3f'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step']))
print('e2e', json.dumps(d['e2e'])[:1400])
print({k:round(v['ms_per_step'],3) for k,v in d['stage_ms'].items()})
PY
tail -3 $out/bench_$tag.err
echo "== bench (first Cholesky kernel)"; PXR_CHOL_V1=1 timeout 700 python bench.py --steps 20 --warmup 5 --cpu-sample-points 0 --no-e2e > $out/bench_${tag}_v1.json 2> $out/bench_${tag}_v1.err; python - <<PY
import json
d=json.load(open('$out/bench_${tag}_v1.json'))
print('value %.1fM ms %.3f steady %.3f'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step']))
print({k:round(v['ms_per_step'],3) for k,v in d['stage_ms'].items()})
PY
echo "== Cholesky time line"
PXR_CHOL_TRACE=$out/chol_trace_$tag.txt timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample-points 0 > /dev/null 2>&1; head -12 $out/chol_trace_$tag.txt; tail -3 $out/chol_trace_$tag.txt
