#!/bin/bash
tag=${1:-r02b}
out=gpurun_out; mkdir -p $out
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_resident.py tests/test_residuals.py tests/test_gpu_upload.py tests/test_gpu_block_mode.py -q -m gpu 2>&1 | tail -25 | tee $out/pytest_new_$tag.log
echo "== bench"; timeout 700 python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<PY
import json
d=json.load(open('$out/bench_$tag.json'))
print('value %.1fM ms %.3f steady %.3f'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step']))
print('e2e', json.dumps(d['e2e'])[:1500])
print({k:round(v['ms_per_step'],3) for k,v in d['stage_ms'].items()})
PY
tail -3 $out/bench_$tag.err
echo "== configs4 N=1"; timeout 600 python bench.py --steps 10 --warmup 3 --workload configs4 --no-e2e --cpu-sample-points 0 > $out/bench_${tag}_configs4_n1.json 2> $out/bench_${tag}_configs4_n1.err; tail -c 600 $out/bench_${tag}_configs4_n1.json; tail -3 $out/bench_${tag}_configs4_n1.err
