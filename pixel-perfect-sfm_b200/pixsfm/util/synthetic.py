"""Seeded synthetic featuremetric-BA / KA scenes (SURVEY.md §8d generator).

Geometry follows the reference's BA test scene (bundle_adjustment/src/bundle_optimizer_test.cc
:73-134): points U(-1,1)^3, SIMPLE_RADIAL cameras f=1200, 1000x1000 images.  For more than
three cameras the cameras sit on a sphere of radius 10 looking at the origin.  Every
observation gets a ps x ps fp16 patch rendered from a smooth per-point C-channel field
F_j(du,dv)[c] = a + b*du + g*dv + h*du*dv, L2-normalised per pixel, plus N(0, noise^2);
patch metadata as FeatureExtractor.tensor_to_fmap produces it (features/extractor.py:190-201).

Not part of the reference's API surface: used by tests/ and bench.py only.
"""
import numpy as np

from .._pixsfm import _capi


def _quat_from_R(R):
    # (w,x,y,z), positive w
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    if q[0] < 0:
        q = -q
    return q / np.linalg.norm(q)


def quat_to_R(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def quat_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def project_simple_radial(params, q, t, X):
    R = quat_to_R(q)
    pc = X @ R.T + t
    u, v = pc[:, 0] / pc[:, 2], pc[:, 1] / pc[:, 2]
    r2 = u * u + v * v
    rad = params[3] * r2
    return np.stack([params[0] * (u + u * rad) + params[1], params[0] * (v + v * rad) + params[2]], 1)


def make_geometry(n_cams, n_points, track_len, seed=0, shared_camera=False):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1, 1, (n_points, 3))
    qs, ts = [], []
    for i in range(n_cams):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        c = 10.0 * d + rng.uniform(-1, 1, 3) * 0.5
        z = -c / np.linalg.norm(c)
        up = np.array([0.0, 1.0, 0.0]) if abs(z[1]) < 0.9 else np.array([1.0, 0.0, 0.0])
        x = np.cross(up, z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z], 0)
        qs.append(_quat_from_R(R)); ts.append(-R @ c)
    qvec, tvec = np.array(qs), np.array(ts)
    n_camera_models = 1 if shared_camera else n_cams
    cam_params = np.tile(np.array([1200.0, 500.0, 500.0, 0.0]), (n_camera_models, 1))
    img_cam = np.zeros(n_cams, np.int32) if shared_camera else np.arange(n_cams, dtype=np.int32)
    track_len = min(track_len, n_cams)
    obs_img = np.empty((n_points, track_len), np.int32)
    for p in range(n_points):
        obs_img[p] = np.sort(rng.choice(n_cams, track_len, replace=False))
    obs_pt = np.repeat(np.arange(n_points, dtype=np.int64), track_len)
    return dict(xyz=xyz, qvec=qvec, tvec=tvec, cam_params=cam_params, img_cam=img_cam,
                obs_img=obs_img.reshape(-1), obs_pt=obs_pt, rng=rng)


def render_patches(xy_true, field_id, n_fields, channels, ps, seed, noise=0.01, dtype=np.float16,
                   image_size=1000, chunk=2048):
    """-> patches [n,ps,ps,C], corners [n,2] int32, scales [n,2]"""
    rng = np.random.default_rng(seed + 1000003)
    a = rng.normal(0, 1.0, (n_fields, channels)).astype(np.float32)
    b = rng.normal(0, 0.15, (n_fields, channels)).astype(np.float32)
    g = rng.normal(0, 0.15, (n_fields, channels)).astype(np.float32)
    h = rng.normal(0, 0.15, (n_fields, channels)).astype(np.float32)
    n = len(xy_true)
    scale = np.ones((n, 2))
    corners = (xy_true * scale - ps / 2.0).astype(np.int32)
    corners = np.clip(corners, [0, 0], np.array([image_size, image_size]) - ps - 1).astype(np.int32)
    patches = np.empty((n, ps, ps, channels), dtype)
    cols = np.arange(ps, dtype=np.float64)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        du = (cols[None, :] + corners[s:e, 0:1] + 0.5 - xy_true[s:e, 0:1]).astype(np.float32)  # [m,ps]
        dv = (cols[None, :] + corners[s:e, 1:2] + 0.5 - xy_true[s:e, 1:2]).astype(np.float32)
        fid = field_id[s:e]
        DU = du[:, None, :, None]; DV = dv[:, :, None, None]
        F = (a[fid][:, None, None, :] + b[fid][:, None, None, :] * DU + g[fid][:, None, None, :] * DV
             + h[fid][:, None, None, :] * DU * DV)
        F /= np.linalg.norm(F, axis=-1, keepdims=True)
        if noise > 0:
            F += rng.normal(0, noise, F.shape).astype(np.float32)
        patches[s:e] = F.astype(dtype)
    return patches, corners, scale


def make_ba_scene(n_cams=6, n_points=60, track_len=4, channels=128, ps=16, seed=0, noise=0.01,
                  dtype=np.float16, rot_sigma_deg=0.02, t_sigma=0.002, pt_sigma=0.005,
                  shared_camera=False, refine_focal=True, refine_pp=False, refine_extra=True,
                  refine_extrinsics=True, with_refs=True):
    """Returns (BAProblem at the perturbed initial state, dict with ground truth)."""
    geo = make_geometry(n_cams, n_points, track_len, seed, shared_camera)
    rng = geo["rng"]
    obs_img, obs_pt = geo["obs_img"], geo["obs_pt"]
    n_obs = len(obs_pt)
    xy_true = np.empty((n_obs, 2))
    for i in range(n_cams):
        m = obs_img == i
        if m.any():
            xy_true[m] = project_simple_radial(geo["cam_params"][geo["img_cam"][i]], geo["qvec"][i],
                                               geo["tvec"][i], geo["xyz"][obs_pt[m]])
    patches, corners, scale = render_patches(xy_true, obs_pt, n_points, channels, ps, seed, noise, dtype)
    # perturb
    q0 = geo["qvec"].copy(); t0 = geo["tvec"].copy(); X0 = geo["xyz"].copy()
    for i in range(n_cams):
        w = rng.normal(0, np.deg2rad(rot_sigma_deg), 3)
        ang = np.linalg.norm(w)
        dq = np.array([1.0, 0, 0, 0]) if ang == 0 else np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])
        q0[i] = quat_mul(dq, q0[i]); q0[i] /= np.linalg.norm(q0[i])
        t0[i] += rng.normal(0, t_sigma, 3)
    X0 += rng.normal(0, pt_sigma, X0.shape)
    n_camera_models = len(geo["cam_params"])
    # masks = BundleOptimizer::ParameterizeCameras with the default_problem_setup gauge
    focal, pp, extra = _capi.CAMERA_PARAM_GROUPS[2]
    mask = 0
    if not refine_focal: mask |= focal
    if not refine_pp: mask |= pp
    if not refine_extra: mask |= extra
    if not (refine_focal or refine_pp or refine_extra): mask = 0xFFFFFFFF
    cam_const_mask = np.full(n_camera_models, mask, np.uint32)
    pose_const = np.zeros(n_cams, np.uint8); tmask = np.zeros(n_cams, np.uint8)
    if refine_extrinsics:
        pose_const[0] = 1            # set_constant_pose(reg_image_ids[0])
        if n_cams > 1: tmask[1] = 1  # set_constant_tvec(reg_image_ids[1], [0])
    else:
        pose_const[:] = 1
    prob = _capi.BAProblem(cam_model=np.full(n_camera_models, 2, np.int32), cam_params=geo["cam_params"],
                           cam_const_mask=cam_const_mask, qvec=q0, tvec=t0, img_cam=geo["img_cam"],
                           pose_const=pose_const, tvec_const_mask=tmask, xyz=X0,
                           point_const=np.zeros(n_points, np.uint8), obs_img=obs_img, obs_pt=obs_pt,
                           patches=patches, corner=corners, scale=scale)
    gt = dict(qvec=geo["qvec"], tvec=geo["tvec"], xyz=geo["xyz"], cam_params=geo["cam_params"],
              xy_true=xy_true)
    return prob, gt


def make_ka_scene(n_images=6, n_tracks=40, track_len=4, channels=128, ps=16, seed=0, noise=0.01,
                  kp_sigma=1.0, dtype=np.float16, extra_edge_prob=0.5):
    """A keypoint-adjustment scene: each track = one 3D point seen in `track_len` images; nodes are
    (image, feature) keypoints detected with N(0,kp_sigma) px error; patches rendered around the
    DETECTED keypoint (corner from the detection, features/extractor.py:192-193) from the true
    field centred on the true projection.  Edges: a spanning chain plus random extra pairs inside a
    track, similarity U(0.5,1).  Returns dict of flat arrays (graph in insertion order)."""
    geo = make_geometry(n_images, n_tracks, track_len, seed, False)
    rng = geo["rng"]
    obs_img, obs_pt = geo["obs_img"], geo["obs_pt"]
    n = len(obs_pt)
    xy_true = np.empty((n, 2))
    for i in range(n_images):
        m = obs_img == i
        if m.any():
            xy_true[m] = project_simple_radial(geo["cam_params"][i], geo["qvec"][i], geo["tvec"][i],
                                               geo["xyz"][obs_pt[m]])
    kps = xy_true + rng.normal(0, kp_sigma, xy_true.shape)
    # patches are cut around the detected keypoint
    rngp = np.random.default_rng(seed + 7)
    scale = np.ones((n, 2))
    corners = np.clip((kps * scale - ps / 2.0).astype(np.int32), [0, 0], np.array([1000, 1000]) - ps - 1).astype(np.int32)
    a = rngp.normal(0, 1.0, (n_tracks, channels)).astype(np.float32)
    b = rngp.normal(0, 0.15, (n_tracks, channels)).astype(np.float32)
    g = rngp.normal(0, 0.15, (n_tracks, channels)).astype(np.float32)
    h = rngp.normal(0, 0.15, (n_tracks, channels)).astype(np.float32)
    cols = np.arange(ps, dtype=np.float64)
    du = (cols[None, :] + corners[:, 0:1] + 0.5 - xy_true[:, 0:1]).astype(np.float32)
    dv = (cols[None, :] + corners[:, 1:2] + 0.5 - xy_true[:, 1:2]).astype(np.float32)
    DU = du[:, None, :, None]; DV = dv[:, :, None, None]
    F = a[obs_pt][:, None, None, :] + b[obs_pt][:, None, None, :] * DU + g[obs_pt][:, None, None, :] * DV \
        + h[obs_pt][:, None, None, :] * DU * DV
    F /= np.linalg.norm(F, axis=-1, keepdims=True)
    F += rngp.normal(0, noise, F.shape).astype(np.float32)
    patches = F.astype(dtype)
    # graph: node i = observation i; image id = obs_img, feature idx = running index per image
    feat_idx = np.zeros(n, np.int32)
    cnt = {}
    for i in range(n):
        feat_idx[i] = cnt.get(int(obs_img[i]), 0); cnt[int(obs_img[i])] = feat_idx[i] + 1
    es, ed, sim = [], [], []
    L = track_len
    for tr in range(n_tracks):
        base = tr * L
        for k in range(L - 1):
            es.append(base + k); ed.append(base + k + 1); sim.append(rng.uniform(0.5, 1.0))
        for k in range(L):
            for m in range(k + 2, L):
                if rng.uniform() < extra_edge_prob:
                    es.append(base + k); ed.append(base + m); sim.append(rng.uniform(0.5, 1.0))
    return dict(keypoints=kps, xy_true=xy_true, node_image=obs_img.astype(np.int32), node_feature=feat_idx,
                edge_src=np.array(es, np.int64), edge_dst=np.array(ed, np.int64), edge_sim=np.array(sim),
                patches=patches, corner=corners, scale=scale, track_gt=obs_pt)
