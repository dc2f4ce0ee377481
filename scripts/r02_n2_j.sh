#!/bin/bash
# 2-GPU consistency check of the final code (fused pair walk in block mode), then a short weak-scaling bench
tag=${1:-r02j_n2}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== consistency check"; timeout 150 $TR --master-port 29611 scripts/check_multi_gpu.py 2>&1 | grep -E "world=|KA sharded|MULTI_GPU|mailbox|Error|error" | tee $out/check_multi_$tag.log
echo "== configs2 weak, N=2"; timeout 100 $TR --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-surface --cpu-sample-points 0 > $out/bench_${tag}.json 2> $out/bench_${tag}.err; tail -c 900 $out/bench_${tag}.json; tail -2 $out/bench_${tag}.err
