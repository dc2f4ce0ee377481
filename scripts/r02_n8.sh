#!/bin/bash
# round-2 8-GPU run: consistency, BASELINE configs[4] (5k cams / 2M pts / 20M obs over 8 GPUs) at N=8 and N=4 (same per-GPU
# shard), strong scaling of configs[2] at N=8 and N=4, and an NCCL call trace of a configs[4] run at N=8.
tag=${1:-r02_n8}
out=gpurun_out; mkdir -p $out
run() { n=$1; port=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
nvidia-smi -L | wc -l
echo "== consistency check N=8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 scripts/check_multi_gpu.py 2>&1 | grep -E "world=|KA sharded|MULTI_GPU|mailbox|Error|error" | tee $out/check_multi_$tag.log
for n in 8 4; do
  echo "== configs4 N=$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 10 --warmup 3 --workload configs4 --no-e2e > $out/bench_${tag}_configs4_n$n.json 2> $out/bench_${tag}_configs4_n$n.err; tail -c 700 $out/bench_${tag}_configs4_n$n.json; tail -2 $out/bench_${tag}_configs4_n$n.err
done
for n in 8 4; do
  echo "== configs2 strong N=$n"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n bench.py --gpus $n --steps 20 --warmup 5 --scaling strong --no-e2e > $out/bench_${tag}_strong_n$n.json 2> $out/bench_${tag}_strong_n$n.err; tail -c 500 $out/bench_${tag}_strong_n$n.json; tail -2 $out/bench_${tag}_strong_n$n.err
done
echo "== NCCL call trace (configs4, N=8, 3 iterations)"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL NCCL_DEBUG_FILE=$out/nccl_trace_${tag}_%h_%p.log timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 8 --steps 3 --warmup 1 --workload configs4 --no-e2e > $out/bench_${tag}_trace.json 2> $out/bench_${tag}_trace.err
for f in $out/nccl_trace_${tag}_*.log; do echo "$f: $(grep -c 'AllReduce' $f) AllReduce lines"; done | head -3
cat $out/nccl_trace_${tag}_*.log | grep -E "NVLS|Init COMPLETE|AllReduce" | head -400 > $out/nccl_trace_${tag}.txt; rm -f $out/nccl_trace_${tag}_*.log
ls -la $out | tail -14
