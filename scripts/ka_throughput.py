"""Keypoint-adjustment throughput on one GPU (python scripts/ka_throughput.py [n_tracks]): packed problems of <= 50
keypoints (keypoint_adjustment/main.py:13-57), one CTA per problem.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixel-perfect-sfm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from ka_util import make_ka_problem
    from pixsfm._pixsfm import _capi, _engine
    n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    t0 = time.time()
    prob, sc, lab = make_ka_problem(n_images=40, n_tracks=n_tracks, track_len=4, channels=128, seed=1, kp_sigma=1.0,
                                    bound=4.0, max_per_problem=50)
    gen = time.time() - t0
    ic = _capi.default_interp(); so = _capi.default_ka_options()
    _engine.ka_run(prob.copy(), ic, so)            # warm-up (context, module load)
    p = prob.copy()
    t0 = time.time()
    s = _engine.ka_run(p, ic, so)
    dt = time.time() - t0
    n_edges = len(prob.edge_src)
    print(json.dumps({"workload": "KA: %d keypoints, %d edges, %d problems, 128-ch fp16 16x16" % (len(prob.keypoints), n_edges, prob.n_problems),
                      "seconds_e2e": dt, "h2d_bytes": s["h2d_bytes"], "initial_cost": s["initial_cost"], "final_cost": s["final_cost"],
                      "edges_per_s_e2e": n_edges / dt, "keypoints_per_s_e2e": len(prob.keypoints) / dt, "scene_generation_s": gen,
                      "message": s.get("message", "")}))


if __name__ == "__main__":
    main()
