#!/bin/bash
# 2 GPUs, final code: configs[4] shard shape (weak) and configs[2] strong scaling
tag=${1:-r02k_n2}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== configs4 N=2"; timeout 70 $TR --master-port 29621 bench.py --gpus 2 --steps 10 --warmup 3 --workload configs4 --no-e2e --no-surface --cpu-sample-points 0 > $out/bench_${tag}_configs4.json 2> $out/bench_${tag}_configs4.err; tail -c 420 $out/bench_${tag}_configs4.json | head -c 300; echo
echo "== configs2 strong N=2"; timeout 50 $TR --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --no-e2e --no-surface --cpu-sample-points 0 > $out/bench_${tag}_strong.json 2> $out/bench_${tag}_strong.err; tail -c 420 $out/bench_${tag}_strong.json | head -c 300; echo
