"""Quick device-time probe of the BA stages on a mid-size numpy-generated scene (not the bench)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixel-perfect-sfm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from pixsfm._pixsfm import _capi, _engine
from pixsfm.util import synthetic
import oracle_lib as O

n_cams, n_pts, tl = int(sys.argv[1]) if len(sys.argv) > 1 else 50, int(sys.argv[2]) if len(sys.argv) > 2 else 6000, 10
t = time.time()
prob, gt = synthetic.make_ba_scene(n_cams=n_cams, n_points=n_pts, track_len=tl, channels=128, seed=1)
print("scene gen %.1fs, n_obs=%d, patches %.2f GB" % (time.time() - t, prob.n_obs, prob.patches.nbytes / 1e9), flush=True)
ic = _capi.default_interp()
t = time.time(); refs, _ = O.refs_compute(prob, ic); prob.refs = refs
print("oracle refs %.1fs" % (time.time() - t), flush=True)
so = _capi.default_ba_options(use_inner_iterations=1, max_num_iterations=5)
h = _engine.BAHandle(prob, ic, so)
names = {0: "K1 fm_eval JAC", 1: "K1 fm_eval COST", 3: "K0 project", 4: "build", 5: "inner", 2: "project+K1+build(+sync)"}
res = {}
for st in (0, 1, 3, 4, 5, 2):
    h.time_stage(st, 2)
    ms = h.time_stage(st, 10)
    res[names[st]] = ms
    print("%-28s %.3f ms  (%.1f M obs/s, %.1f GB/s @4736B/obs)" % (names[st], ms, prob.n_obs / ms / 1e3, prob.n_obs * 4736 / ms / 1e6), flush=True)
t = time.time(); s = h.solve(); dt = time.time() - t
print("solve: %d its in %.3fs (%.2f ms/it), cost %.6f -> %.6f, launches %d" % (s["num_iterations"] - 1, dt, 1e3 * dt / max(1, s["num_iterations"] - 1), s["initial_cost"], s["final_cost"], s["kernel_launches"]))
for it in s["iterations"]: print("  it %d cost %.9f ok %d t %.2f ms" % (it["iteration"], it["cost"], it["step_is_successful"], it["iteration_time_s"] * 1e3))
so2 = _capi.default_ba_options(use_inner_iterations=0, max_num_iterations=5)
h2 = _engine.BAHandle(prob.copy(), ic, so2)
t = time.time(); s2 = h2.solve(); dt = time.time() - t
print("solve (no inner): %.2f ms/it" % (1e3 * dt / max(1, s2["num_iterations"] - 1)))
for it in s2["iterations"]: print("  it %d cost %.9f ok %d t %.2f ms" % (it["iteration"], it["cost"], it["step_is_successful"], it["iteration_time_s"] * 1e3))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "quick_stage_times.json"), "w"), indent=1)
