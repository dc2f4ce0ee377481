#!/bin/bash
# final-code check on 2 GPUs: consistency against the oracle, the default bench (weak, with the window-resident e2e on both ranks)
tag=${1:-r02_n2_final}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== consistency check"; timeout 600 $TR --master-port 29611 scripts/check_multi_gpu.py 2>&1 | grep -E "world=|KA sharded|MULTI_GPU|mailbox|Error|error" | tee $out/check_multi_$tag.log
echo "== configs2 weak (default bench, e2e included)"; timeout 900 $TR --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_${tag}.json 2> $out/bench_${tag}.err; python - <<PY
import json
d=json.loads(open('$out/bench_${tag}.json').read().strip().splitlines()[-1])
print('value %.1fM ms %.3f steady %.3f coll/it %s'%(d['value']/1e6,d['ms_per_step'],d['steady_state']['ms_per_step'],d['multi_gpu']['nccl_collectives_per_lm_iteration']))
print('e2e', json.dumps(d['e2e'])[:1200])
PY
tail -3 $out/bench_${tag}.err
