#!/bin/bash
# compute-sanitizer memcheck over a representative slice of the GPU tests (run on the GPU box):
#   scripts/memcheck.sh > gpurun_out/memcheck.log 2>&1
# Every kernel family is touched once: K0/K1 (all dtypes), build, Schur, persistent Cholesky, PCG dense + block-sparse,
# inner iterations, reference extraction, cost maps, KA (edge and query mode).
set -u
run() {
  echo "=== $*"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 5 python -m pytest "$@" -q -m gpu -x 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|=========.*(Invalid|Error)" | head -12
}
run tests/test_gpu_ba_parity.py -k "residual_blocks or full_solve or tile_dag or block_sparse or iterative_schur_pcg"
run tests/test_gpu_refs.py
run tests/test_gpu_costmaps.py -k "extraction_matches or fused or full_solve"
run tests/test_gpu_ka.py
run tests/test_gpu_edge_cases.py -k "triangulation or single_free or clamped"
