"""BASELINE configs[3] at shape on one GPU: ETH3D-courtyard-like triangulation refinement — 38 images (the courtyard scene's
size; the reference tree only names the scene, eval/eth3d/config.py:7-8), multi-level features processed coarse to fine
(S2DNet levels at 1/16, 1/4 and full resolution), strategy `costmaps`: per level reference extraction + cost maps
(pxr_costmaps_compute, the 3-channel maps stay on the device) then cost-map BA with poses and intrinsics fixed, 10
iterations, inner iterations on (configs/pixsfm_eth3d.yaml; bundle_adjustment/main.py:218-286).

    python scripts/bench_configs3.py [n_points] [--cpu]

Prints one JSON line: device time per level (CUDA events are inside the library; here wall time around synchronous calls
with the features already resident), observations/s over the three levels, and with --cpu the oracle (CPU restatement of
the reference path, test infrastructure) on a bounded sample of the same scene."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixel-perfect-sfm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def scene(n_cams, n_points, track, seed=0):
    from pixsfm.util import synthetic
    geo = synthetic.make_geometry(n_cams, n_points, track, seed, False)
    obs_img, obs_pt = geo["obs_img"], geo["obs_pt"]
    xy = np.empty((len(obs_pt), 2))
    for i in range(n_cams):
        m = obs_img == i
        if m.any():
            xy[m] = synthetic.project_simple_radial(geo["cam_params"][geo["img_cam"][i]], geo["qvec"][i], geo["tvec"][i], geo["xyz"][obs_pt[m]])
    rng = np.random.default_rng(seed + 5)
    X0 = geo["xyz"] + rng.normal(0, 0.02, geo["xyz"].shape)          # triangulation-quality points: a few px of reprojection error
    return geo, xy, X0


def level_problem(geo, xy, X0, scale, patches, on_device, ps, channels, sel=None):
    from pixsfm._pixsfm import _capi
    n_cams = len(geo["qvec"]); n_models = len(geo["cam_params"])
    n_obs = len(geo["obs_pt"]) if sel is None else sel[1]
    n_pts = len(X0) if sel is None else sel[0]
    size = max(int(round(1000 * scale)), ps + 2)
    corners = np.clip((xy[:n_obs] * scale - ps / 2.0).astype(np.int32), [0, 0], np.array([size, size]) - ps - 1).astype(np.int32)
    kw = dict(cam_model=np.full(n_models, 2, np.int32), cam_params=geo["cam_params"], cam_const_mask=np.full(n_models, 0xFFFFFFFF, np.uint32),
              qvec=geo["qvec"], tvec=geo["tvec"], img_cam=geo["img_cam"], pose_const=np.ones(n_cams, np.uint8),
              tvec_const_mask=np.zeros(n_cams, np.uint8), xyz=X0[:n_pts], point_const=np.zeros(n_pts, np.uint8),
              obs_img=geo["obs_img"][:n_obs], obs_pt=geo["obs_pt"][:n_obs], corner=corners, scale=np.full((n_obs, 2), scale))
    if on_device:
        return _capi.BAProblem(patches=patches, patches_on_device=True, patch_shape=(n_obs, ps, ps, channels), patch_dtype=0, **kw), corners
    return _capi.BAProblem(patches=patches, **kw), corners


def main():
    from pixsfm._pixsfm import _capi, _engine
    n_points = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 40000
    with_cpu = "--cpu" in sys.argv
    n_cams, track, ps, C_ = 38, 5, 16, 128
    scales = (0.0625, 0.25, 1.0)                     # coarse to fine (the adjuster walks the levels in reverse index order)
    geo, xy, X0 = scene(n_cams, n_points, track)
    n_obs = len(geo["obs_pt"])
    ctx = _capi.default_context()
    ic = _capi.default_interp()
    ic_cm = _capi.default_interp(); ic_cm.l2_normalize = 0
    so = _capi.default_ba_options(max_num_iterations=10)
    cfg = _capi.default_costmap_config()
    # warm-up outside the timed region: module load, pinned staging buffers, the first launches of every kernel
    n_w = min(n_points, 500) * track
    d_w = _engine.synth_patches_device(n_w, ps, C_, xy[:n_w] - 0.5 - np.clip((xy[:n_w] - ps / 2.0).astype(np.int32), 0, 1000 - ps - 1),
                                       geo["obs_pt"][:n_w], seed=7, noise=0.01, ctx=ctx)
    pw, _ = level_problem(geo, xy, X0, 1.0, d_w, True, ps, C_, sel=(n_w // track, n_w))
    ow = _engine.costmaps_compute(pw, ic, cfg, to_host=False, to_device=True, ctx=ctx)
    cw = _capi.BAProblem(cam_model=pw.cam_model, cam_params=pw.cam_params, cam_const_mask=pw.cam_const_mask, qvec=pw.qvec, tvec=pw.tvec,
                         img_cam=pw.img_cam, pose_const=pw.pose_const, tvec_const_mask=pw.tvec_const_mask, xyz=pw.xyz.copy(),
                         point_const=pw.point_const, obs_img=pw.obs_img, obs_pt=pw.obs_pt, patches=ow["device_ptr"], corner=pw.corner,
                         scale=pw.scale, patches_on_device=True, patch_shape=(n_w, ps, ps, 3), patch_dtype=0)
    _engine.ba_run(cw, ic_cm, so, ctx=ctx)
    _engine.device_free(ow["device_ptr"], ctx); _engine.device_free(d_w, ctx)
    levels = []
    X = X0.copy()
    total_obs_iters = 0
    t_all = 0.0
    for lvl, sc in enumerate(scales):
        size = max(int(round(1000 * sc)), ps + 2)
        corners = np.clip((xy * sc - ps / 2.0).astype(np.int32), [0, 0], np.array([size, size]) - ps - 1).astype(np.int32)
        uv0 = xy * sc - 0.5 - corners
        d_feat = _engine.synth_patches_device(n_obs, ps, C_, uv0, geo["obs_pt"], seed=77 + lvl, noise=0.01, ctx=ctx)
        prob, _ = level_problem(geo, xy, X, sc, d_feat, True, ps, C_)
        ctx.sync()
        t0 = time.time()
        out = _engine.costmaps_compute(prob, ic, cfg, to_host=False, to_device=True, ctx=ctx)      # references + cost maps, on the device
        ctx.sync()
        t1 = time.time()
        cm = _capi.BAProblem(cam_model=prob.cam_model, cam_params=prob.cam_params, cam_const_mask=prob.cam_const_mask, qvec=prob.qvec,
                             tvec=prob.tvec, img_cam=prob.img_cam, pose_const=prob.pose_const, tvec_const_mask=prob.tvec_const_mask,
                             xyz=X, point_const=prob.point_const, obs_img=prob.obs_img, obs_pt=prob.obs_pt, patches=out["device_ptr"],
                             corner=prob.corner, scale=prob.scale, patches_on_device=True, patch_shape=(n_obs, ps, ps, 3), patch_dtype=0)
        s = _engine.ba_run(cm, ic_cm, so, ctx=ctx)
        ctx.sync()
        t2 = time.time()
        X = cm.xyz.copy()
        its = max(1, s["num_iterations"] - 1)
        total_obs_iters += n_obs * its
        t_all += t2 - t0
        levels.append({"scale": sc, "costmap_extraction_s": t1 - t0, "costmap_ba_s": t2 - t1, "lm_iterations": its,
                       "initial_cost": s["initial_cost"], "final_cost": s["final_cost"], "lm_loop_s": s["solve_time_s"]})
        _engine.device_free(out["device_ptr"], ctx); _engine.device_free(d_feat, ctx)
    line = {"workload": "configs[3] shape: %d images / %d points / %d observations per level, 3 levels of 128-ch fp16 16x16 patches, costmap strategy, "
                        "poses + intrinsics fixed, 10 iterations per level" % (n_cams, n_points, n_obs),
            "seconds_all_levels": t_all, "observations_per_s": total_obs_iters / t_all, "levels": levels,
            "features": "synthetic, generated on the device per level (a feature store that keeps the CNN output on the GPU); cost maps stay on the device"}
    if with_cpu:
        import oracle_lib as O
        n_p = min(n_points, 2000); n_o = n_p * track
        cpu = 0.0
        Xc = X0.copy()
        for lvl, sc in enumerate(scales):
            size = max(int(round(1000 * sc)), ps + 2)
            from pixsfm.util import synthetic
            patches, corners, _ = synthetic.render_patches(xy[:n_o] * sc, geo["obs_pt"][:n_o], n_p, C_, ps, 77 + lvl, 0.01, np.float16, image_size=size)
            p, _ = level_problem(geo, xy, Xc, sc, patches, False, ps, C_, sel=(n_p, n_o))
            p.corner[:] = corners
            t0 = time.time()
            p.refs = O.refs_compute(p, ic, iters=100)[0]
            cmaps = O.costmaps_compute(p)
            q = p.with_patches(cmaps)
            O.ba_solve(q, ic_cm, so)
            cpu += time.time() - t0
            Xc[:n_p] = q.xyz
        line["cpu_oracle"] = {"sample": "%d points / %d observations per level" % (n_p, n_o), "seconds_all_levels": cpu,
                              "threads": int(O.lib().orc_num_threads())}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
