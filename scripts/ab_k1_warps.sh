#!/bin/bash
# A/B of the K1 warps-per-CTA build knob on the GPU box (nvcc is in the image). Usage: scripts/ab_k1_warps.sh 16 20 24
set -u
mkdir -p gpurun_out
for W in "$@"; do
  rm -f pixel-perfect-sfm_b200/csrc/pxr_ba.o pixel-perfect-sfm_b200/csrc/pxr_inner.o pixel-perfect-sfm_b200/csrc/pxr_refs.o
  PXR_FM_WARPS=$W python pixel-perfect-sfm_b200/build.py > gpurun_out/build_w$W.log 2>&1 || { echo "build W=$W failed"; tail -5 gpurun_out/build_w$W.log; continue; }
  python bench.py --steps 5 --warmup 3 --no-e2e --cpu-sample-points 200 > gpurun_out/ab_w$W.json 2> gpurun_out/ab_w$W.err || { echo "bench W=$W failed"; tail -5 gpurun_out/ab_w$W.err; continue; }
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_w$W.json"))
s=d["stage_ms"]
print("W=$W ms/step=%.3f K1cost=%.3f K1jac=%.3f frac=%.3f" % (d["ms_per_step"], s["K1 cost-only"]["ms_per_step"], s["K1 residual/Jacobian"]["ms_per_step"], d["roofline"]["frac"]))
PY
done
