#!/bin/bash
tag=${1:-r02e}
out=gpurun_out; mkdir -p $out
echo "== tests touching inner iterations / residency / determinism"; timeout 900 python -m pytest tests/test_gpu_ba_parity.py tests/test_gpu_resident.py tests/test_gpu_deterministic.py tests/test_gpu_edge_cases.py tests/test_gpu_costmaps.py tests/test_gpu_block_mode.py tests/test_gpu_full_size.py -q -m gpu 2>&1 | tail -8
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-surface > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<PY
import json
d=json.load(open('$out/bench_$tag.json'))
print('value %.1fM ms %.3f steady %.3f'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step']))
print('e2e', json.dumps({k:d['e2e'][k] for k in ('value','seconds','library_seconds','lm_loop_seconds','first_call_seconds')}), 'full', d['e2e']['full_upload']['seconds'])
print({k:round(v['ms_per_step'],3) for k,v in d['stage_ms'].items()})
PY
tail -2 $out/bench_$tag.err
