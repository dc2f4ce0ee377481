#!/bin/bash
# round-2 state check on one B200: smoke, the GPU suite (no -x: one failure must not hide the rest), both bench arms
tag=${1:-r02a}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -1; nproc; free -g | head -2
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -40 | tee $out/pytest_gpu_$tag.log
echo "== bench"; timeout 700 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 3000 $out/bench_$tag.json; tail -5 $out/bench_$tag.err
echo "== bench --impl reference"; timeout 700 python bench.py --impl reference > $out/bench_${tag}_reference.json 2> $out/bench_${tag}_reference.err; tail -c 1500 $out/bench_${tag}_reference.json; tail -5 $out/bench_${tag}_reference.err
