#!/bin/bash
tag=${1:-r02f}
out=gpurun_out; mkdir -p $out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest -m gpu -x (as the driver runs it)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $out/pytest_gpu_$tag.log
echo "== configs3 shape"; timeout 600 python scripts/bench_configs3.py 40000 --cpu 2>$out/bench_configs3_$tag.err | tail -1 | tee $out/bench_configs3_$tag.json | cut -c1-1500; tail -2 $out/bench_configs3_$tag.err
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; python - <<PY
import json
d=json.load(open('$out/bench_$tag.json'))
print('value %.1fM ms %.3f steady %.3f'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step']))
print('e2e', json.dumps({k:d['e2e'][k] for k in ('value','seconds','library_seconds','lm_loop_seconds','first_call_seconds','observations_refetched')}), 'full', d['e2e']['full_upload']['seconds'])
print('surface', json.dumps(d.get('e2e_reference_surface'))[:420])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
tail -2 $out/bench_$tag.err
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_${tag}_reference.json 2> $out/bench_${tag}_reference.err; python - <<PY
import json
d=json.load(open('$out/bench_${tag}_reference.json')); print('reference arm', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:120])
PY
