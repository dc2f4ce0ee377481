#!/bin/bash
# round-2 multi-GPU checks on N GPUs of one box:  bash scripts/r02_multi.sh N tag [full]
#   consistency of the point-sharded BA against the single-process oracle, weak / strong scaling of BASELINE configs[2],
#   BASELINE configs[4] (5k cameras, 250k points / 2.5M observations per GPU), and an NCCL call trace of a short run.
N=${1:-2}; tag=${2:-r02_n$N}; full=${3:-}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | head -8; nvidia-smi topo -m 2>/dev/null | head -12
echo "== consistency check"; timeout 600 $TR --master-port 29611 scripts/check_multi_gpu.py 2>&1 | grep -v -E '^\s*$|W[0-9]{4}' | tail -25 | tee $out/check_multi_$tag.log
echo "== configs2 weak"; timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_${tag}_weak.json 2> $out/bench_${tag}_weak.err; tail -c 1200 $out/bench_${tag}_weak.json; tail -3 $out/bench_${tag}_weak.err
echo "== configs2 strong"; timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --no-e2e > $out/bench_${tag}_strong.json 2> $out/bench_${tag}_strong.err; tail -c 600 $out/bench_${tag}_strong.json; tail -3 $out/bench_${tag}_strong.err
PTS=${CONFIGS4_POINTS:-250000}
echo "== configs4 ($PTS points per GPU)"; timeout 900 $TR --master-port 29614 bench.py --gpus $N --steps 10 --warmup 3 --workload configs4 --points $PTS --no-e2e > $out/bench_${tag}_configs4.json 2> $out/bench_${tag}_configs4.err; tail -c 1500 $out/bench_${tag}_configs4.json; tail -3 $out/bench_${tag}_configs4.err
echo "== NCCL call trace (configs2 weak, 5 iterations)"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 $TR --master-port 29615 bench.py --gpus $N --steps 5 --warmup 1 --no-e2e > $out/bench_${tag}_trace.json 2> $out/nccl_trace_$tag.log
grep -c "AllReduce" $out/nccl_trace_$tag.log; grep -E "NVLS|nranks|Connected all" $out/nccl_trace_$tag.log | head -6
ls -la $out | tail -12
