"""The reference-facing Python layer end to end WITHOUT a device: the device entry points of `_engine` are replaced by
the oracle (same problem IR in, same summary/arrays out), and the mirror-level GPU tests are run as they are.  What
this covers is the host logic above the C-ABI — problem construction from reconstructions / graphs / databases /
model files, option handling, write-back — which is the same code whichever side solves the IR.  (The product never
does this; it is test plumbing, like everything that touches oracle/.)"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from pixsfm._pixsfm import _capi, _engine


def _ka_run(problem, interp=None, options=None, ctx=None):
    interp = interp or _capi.default_interp()
    options = options or _capi.default_ka_options()
    d = problem.desc()
    s = _capi.make_summary(0)
    O.lib().orc_ka_solve(C.byref(d), C.byref(interp), C.byref(options), C.byref(s))
    out = _capi.summary_to_dict(s)
    out["num_residual_blocks"] = len(problem.edge_src)
    out["num_residuals"] = len(problem.edge_src) * problem.channels
    return out


def _ba_run(problem, interp=None, options=None, ctx=None, capacity=512):
    out = O.ba_solve(problem, interp or _capi.default_interp(), options or _capi.default_ba_options())
    out.setdefault("num_residual_blocks", problem.n_obs)
    return out


def _refs_compute(problem, interp=None, loss_type=1, loss_scale=0.25, iters=100, ctx=None):
    return O.refs_compute(problem, interp or _capi.default_interp(), loss_type, loss_scale, iters)


def _costmaps_compute(problem, interp=None, cfg=None, refs=None, to_host=True, to_device=False, ctx=None):
    interp = interp or _capi.default_interp()
    cfg = cfg or _capi.default_costmap_config()
    src = np.full(len(problem.xyz), -1, np.int64)
    if cfg.compute_refs:
        refs, src = O.refs_compute(problem, interp, cfg.ref_loss_type, cfg.ref_loss_scale, cfg.ref_iters)
    p2 = problem.copy()
    p2.refs = np.ascontiguousarray(refs if refs is not None else problem.refs, np.float64)
    out = O.costmaps_compute(p2, cfg.loss_type, cfg.loss_scale, bool(cfg.as_gradientfield), bool(cfg.apply_sqrt))
    return {"costmaps": out, "device_ptr": None, "refs": p2.refs, "src_obs": src, "summary": {}}


def _obs_descriptors(problem, interp=None, ctx=None):
    interp = interp or _capi.default_interp()
    xy = O.ba_evaluate(problem.copy() if problem.refs is not None else _with_zero_refs(problem), interp,
                       _capi.default_ba_options())["xy"]
    out = np.zeros((problem.n_obs, problem.channels))
    for o in range(problem.n_obs):
        pi = int(problem.obs_patch[o]) if problem.obs_patch is not None and len(problem.obs_patch) else o
        uv = (xy[o] * problem.scale[pi] - 0.5 - problem.corner[pi]) * problem.upsampling_factor
        out[o] = O.pixel_interp(_patch_of(problem, pi), uv[1], uv[0], bool(interp.l2_normalize), bool(interp.use_float_simd))[0]
    return out


def _with_zero_refs(problem):
    q = problem.copy()
    q.refs = np.zeros((len(problem.xyz), problem.channels))
    return q


def _patch_of(problem, pi):
    if problem.patch_blocks is not None:
        starts = np.concatenate([[0], np.cumsum([b.shape[0] for b in problem.patch_blocks])])
        b = int(np.searchsorted(starts, pi, side="right") - 1)
        return np.asarray(problem.patch_blocks[b][pi - starts[b]])
    return problem.patches[pi]


def _interpolate_descriptors(fmap, patch_idxs, xys, interpolation_config=None):
    from pixsfm._pixsfm import _localization as L
    from pixsfm._pixsfm._base import InterpolationConfig
    interp = interpolation_config if isinstance(interpolation_config, InterpolationConfig) else InterpolationConfig(interpolation_config or {})
    patches = L._host_patches(fmap)
    xys = np.asarray(xys, np.float64).reshape(-1, 2)
    out = np.zeros((len(xys), fmap.channels))
    for k, i in enumerate(patch_idxs):
        li = fmap.local_index(i)
        uv = xys[k] * np.asarray(fmap.scale, np.float64) - 0.5 - fmap.corners[li]
        out[k] = O.pixel_interp(patches[li], uv[1], uv[0], bool(interp.l2_normalize), bool(interp.use_float_simd))[0]
    return out


@pytest.fixture
def oracle_engine(monkeypatch):
    from pixsfm._pixsfm import _localization as L
    monkeypatch.setattr(_engine, "obs_descriptors", _obs_descriptors)
    monkeypatch.setattr(L, "interpolate_descriptors", _interpolate_descriptors)
    monkeypatch.setattr(_engine, "ka_run", _ka_run)
    monkeypatch.setattr(_engine, "ba_run", _ba_run)
    monkeypatch.setattr(_engine, "refs_compute", _refs_compute)
    monkeypatch.setattr(_engine, "costmaps_compute", _costmaps_compute)


def test_bundle_adjuster_multilevel(oracle_engine):
    import test_gpu_mirror as T
    T.test_bundle_adjuster_refine_multilevel_matches_oracle()


def test_keypoint_adjusters(oracle_engine):
    import test_gpu_mirror as T
    T.test_keypoint_adjuster_refine_multilevel_moves_keypoints_towards_truth()
    T.test_topological_reference_keypoint_adjuster_matches_flat_oracle()


def test_dense_feature_maps(oracle_engine):
    import test_gpu_mirror as T
    T.test_dense_feature_maps_match_oracle()


def test_colmap_database_and_model_directory_drivers(oracle_engine, tmp_path):
    import test_gpu_mirror as T
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    T.test_keypoint_adjustment_from_a_colmap_database(tmp_path / "a")
    T.test_pixsfm_refines_a_colmap_model_directory(tmp_path / "b")


def test_costmap_bundle_adjuster(oracle_engine):
    import test_gpu_costmaps as T
    T.test_costmap_bundle_adjuster_strategy_end_to_end()


@pytest.mark.parametrize("arg", [0, 1])
def test_query_adjusters(oracle_engine, arg):
    import test_gpu_localization as T
    params = [m.args[1] for m in getattr(T.test_query_keypoint_adjuster_matches_oracle, "pytestmark", []) if m.name == "parametrize"]
    T.test_query_keypoint_adjuster_matches_oracle(params[0][arg] if params else arg)
    params = [m.args[1] for m in getattr(T.test_query_bundle_adjuster_matches_oracle, "pytestmark", []) if m.name == "parametrize"]
    T.test_query_bundle_adjuster_matches_oracle(params[0][arg] if params else bool(arg))


def test_localizer_pipeline(oracle_engine):
    import test_gpu_localization as T
    T.test_descriptor_interpolation_and_nearest_references()
    T.test_query_localizer_runs_qka_pnp_qba()


def test_query_helpers_and_stacked_correspondences(oracle_engine):
    """localization/main.py helpers: the feature-inlier test only judges single-descriptor references, and stacking
    the correspondences of one keypoint solves it once against all of its targets."""
    import test_gpu_localization as T
    from pixsfm import localization as loc_pkg
    from pixsfm._pixsfm import _localization as L
    from pixsfm.localization import main as M
    rec, fm, refs, qid, fmap, p2D_idxs, p3D_ids, kps = T._scene()
    desc = L.interpolate_descriptors(fmap, p2D_idxs, kps, {})
    near = [desc[k].copy() for k in range(len(kps))]
    far = [-desc[k] for k in range(len(kps))]
    mixed = [near[0], far[1], refs[p3D_ids[2]]] + near[3:]
    got = M.find_feature_inliers(kps, fmap, mixed, {}, thresh=0.5, point2D_idxs=p2D_idxs)
    assert got[0] is True and got[1] is False and got[2] is True and all(got[3:])       # a Reference object is not judged
    assert M.find_feature_inliers(kps, fmap, far, {}, thresh=-1, point2D_idxs=p2D_idxs) == [True] * len(kps)
    # stacked QKA: duplicate every correspondence (same keypoint, two targets) -> the duplicates move together and the
    # result equals solving the distinct keypoints once against both targets
    rng = np.random.default_rng(1)
    start = kps + rng.normal(0, 0.7, kps.shape)
    t1 = [np.asarray(refs[p].descriptor, np.float64).reshape(-1) for p in p3D_ids]
    t2 = [np.asarray(refs[p].observations[0], np.float64).reshape(-1) for p in p3D_ids]
    dup_kps = np.concatenate([start, start])
    dup_idx = list(p2D_idxs) + list(p2D_idxs)
    qka = loc_pkg.QueryKeypointAdjuster({"stack_correspondences": True})
    qka.refine(dup_kps, fmap, t1 + t2, point2D_idxs=dup_idx)
    n = len(kps)
    assert np.array_equal(dup_kps[:n], dup_kps[n:]) and np.abs(dup_kps[:n] - start).max() > 1e-3
    direct = start.copy()
    L.QueryKeypointOptimizer(qka.solver.options, qka.solver.interp).run(direct, fmap, [[a, b] for a, b in zip(t1, t2)],
                                                                        patch_idxs=list(p2D_idxs))
    assert np.abs(direct - dup_kps[:n]).max() < 1e-9
    with pytest.raises(ValueError, match="np.ndarray reference"):
        qka.refine(dup_kps, fmap, [refs[p] for p in p3D_ids] * 2, point2D_idxs=dup_idx)
    with pytest.raises(ValueError, match="point2D_idxs must not be None"):
        qka.refine(dup_kps, fmap, t1 + t2)


def test_two_phase_bundle_optimizer_surface(oracle_engine):
    """set_up / problem / solve_problem / reset, as bound by bundle_adjustment/bindings.cc:36-51"""
    import copy
    from pixsfm import bundle_adjustment as ba_pkg, features
    from pixsfm._pixsfm import _bundle_adjustment as ba
    from recon_util import make_reconstruction
    rec, fm, _, _ = make_reconstruction(n_cams=5, n_points=40, track_len=3, channels=16, seed=31)
    rec2 = copy.deepcopy(rec)
    setup = ba_pkg.default_problem_setup(rec)
    labels = ba_pkg.find_problem_labels(rec, 10)
    refs = ba.ReferenceExtractor({}, {}).run(labels, rec, fm.fset(0))
    fview = features.FeatureView(fm.fset(0), rec)
    opt = ba.FeatureReferenceBundleOptimizer({"solver": {"max_num_iterations": 5}}, setup, {})
    with pytest.raises(ValueError, match="set_up"):
        opt.solve_problem()
    assert opt.problem is None and opt.set_up(rec, fview, refs) is True
    prob = opt.problem
    assert prob.n_obs == rec.num_observations() and len(prob.xyz) == len(rec.points3D)
    with pytest.raises(ValueError, match="multiple times"):
        opt.set_up(rec, fview, refs)
    before = {i: im.tvec.copy() for i, im in rec.images.items()}
    assert opt.solve_problem() is True and opt.summary().final_cost < opt.summary().initial_cost
    assert any(not np.array_equal(rec.images[i].tvec, before[i]) for i in before)          # written back
    # one-shot run on the untouched copy gives the same result
    one = ba.FeatureReferenceBundleOptimizer({"solver": {"max_num_iterations": 5}}, ba_pkg.default_problem_setup(rec2), {})
    assert one.run(rec2, features.FeatureView(fm.fset(0), rec2), refs) is True
    for i in rec.images:
        assert np.abs(rec.images[i].tvec - rec2.images[i].tvec).max() < 1e-9      # (the oracle sums with OpenMP)
    opt.reset()
    assert opt.problem is None and opt.summary() is None and opt.set_up(rec, fview, refs) is True


def test_keypoint_optimizer_run_subset(oracle_engine):
    """run_subset (keypoint_adjustment/bindings.cc:17-23): the nodes of two tracks only; everything else stays put, and
    the moved keypoints equal what a run over labels that isolate those tracks gives"""
    from pixsfm import base, features
    from pixsfm._pixsfm import _keypoint_adjustment as ka
    from pixsfm.util import synthetic
    sc = synthetic.make_ka_scene(n_images=5, n_tracks=12, track_len=4, channels=16, seed=21, kp_sigma=0.8)
    g = base.Graph()
    names = ["im%d" % i for i in range(5)]
    keypoints = {names[i]: np.ascontiguousarray(sc["keypoints"][sc["node_image"] == i]) for i in range(5)}
    for n in range(len(sc["node_image"])):
        g.add_node(names[sc["node_image"][n]], int(sc["node_feature"][n]))
    for e in range(len(sc["edge_src"])):
        g.add_edge(g.nodes[sc["edge_src"][e]], g.nodes[sc["edge_dst"][e]], sc["edge_sim"][e])
    fset = features.FeatureSet()
    for i in range(5):
        m = np.where(sc["node_image"] == i)[0]
        fset.emplace(names[i], features.FeatureMap(np.ascontiguousarray(sc["patches"][m]), sc["node_feature"][m].tolist(),
                                                   sc["corner"][m], {"scale": sc["scale"][m[0]], "is_sparse": True}))
    tl = base.compute_track_labels(g)
    roots = base.compute_root_labels(g, tl, base.compute_score_labels(g, tl))
    setup = ka.KeypointAdjustmentSetup(); setup.set_masked_nodes_constant(g, roots)
    chosen = [n for n in range(len(tl)) if tl[n] in (tl[0], tl[7])]
    kp_sub = {k: v.copy() for k, v in keypoints.items()}
    summary = ka.FeatureMetricKeypointOptimizer({}, setup, {}).run_subset(chosen, kp_sub, g, tl, roots, fset)
    assert summary.final_cost < summary.initial_cost and summary.num_residual_blocks > 0
    moved = {(g.image_id_to_name[g.nodes[n].image_id], g.nodes[n].feature_idx) for n in chosen if not roots[n]}
    for name in names:
        for f in range(len(keypoints[name])):
            same = np.array_equal(kp_sub[name][f], keypoints[name][f])
            assert same != ((name, f) in moved)
    with pytest.raises(ValueError):
        ka.FeatureMetricKeypointOptimizer({}, setup, {}).run_subset([10 ** 6], kp_sub, g, tl, roots, fset)


def _interpolate_patches(patches, corners, scales, item_patch, xys, interp=None, upsampling_factor=1.0, ctx=None):
    interp = interp or _capi.default_interp()
    out = np.zeros((len(xys), np.shape(patches)[3]))
    for k, (pi, xy) in enumerate(zip(item_patch, xys)):
        uv = (np.asarray(xy, np.float64) * np.asarray(scales[pi], np.float64) - 0.5 - np.asarray(corners[pi])) * upsampling_factor
        out[k] = O.pixel_interp(np.asarray(patches)[pi], uv[1], uv[0], bool(interp.l2_normalize), bool(interp.use_float_simd))[0]
    return out


def test_patch_interpolator_surface(monkeypatch):
    """features.PatchInterpolator: image vs local coordinates, nodes, config handling (oracle as the evaluator)"""
    from pixsfm import features
    monkeypatch.setattr(_engine, "interpolate_patches", _interpolate_patches)
    rng = np.random.default_rng(4)
    data = rng.normal(size=(16, 16, 8)).astype(np.float16)
    patch = features.FeaturePatch(data, (100, 40), (0.5, 0.25))
    pi = features.PatchInterpolator({"l2_normalize": True})
    xy = np.array([213.3, 188.9])
    uv = patch.to_pixel_coordinates(xy)
    assert np.allclose(uv, [213.3 * 0.5 - 0.5 - 100, 188.9 * 0.25 - 0.5 - 40])
    f = pi.interpolate(patch, xy)
    assert f.shape == (1, 8) and abs(np.linalg.norm(f) - 1) < 1e-12
    assert np.abs(f - O.pixel_interp(data, uv[1], uv[0])[0]).max() < 1e-15
    assert np.array_equal(pi.interpolate_nodes(patch, xy), f)
    assert np.abs(pi.interpolate_local(patch, uv) - f).max() < 1e-12
    raw = features.PatchInterpolator({"l2_normalize": False}).interpolate_local(patch, [3.0, 5.0])
    assert np.abs(raw[0] - data[5, 3].astype(np.float64)).max() < 1e-6         # at a pixel centre the spline interpolates
    with pytest.raises(ValueError):
        features.PatchInterpolator({"nodes": [[0, 0], [1, 0]]}).interpolate(patch, xy)


def test_costmaps_from_dense_feature_maps(oracle_engine):
    """The cost-map strategy on DENSE feature maps (costmap_extractor.h:186-224, 398-428): a dense_cut_size window per
    observation around its reprojection (corner rule: FeaturePatch::ToCorner), references as from the dense map itself."""
    import copy
    from pixsfm import bundle_adjustment as ba_pkg, features
    from pixsfm._pixsfm import _bundle_adjustment as ba
    from pixsfm.util import cameras
    from recon_util import make_reconstruction
    rec, fm_sparse, _, _ = make_reconstruction(n_cams=5, n_points=40, track_len=3, channels=16, seed=41)
    rng = np.random.default_rng(7)
    H = W = 72
    fm = features.FeatureManager([16], np.float16)
    for name in fm_sparse.fset(0).keys():
        base_f = rng.normal(size=(H // 8 + 2, W // 8 + 2, 16))          # a smooth random field
        yy, xx = np.meshgrid(np.linspace(1, H // 8, H), np.linspace(1, W // 8, W), indexing="ij")
        y0, x0 = yy.astype(int), xx.astype(int)
        fy, fx = (yy - y0)[..., None], (xx - x0)[..., None]
        dense = ((1 - fy) * (1 - fx) * base_f[y0, x0] + (1 - fy) * fx * base_f[y0, x0 + 1]
                 + fy * (1 - fx) * base_f[y0 + 1, x0] + fy * fx * base_f[y0 + 1, x0 + 1]).astype(np.float16)
        fm.fset(0).emplace(name, features.FeatureMap(np.ascontiguousarray(dense[None]), [features.kDenseId], np.zeros((1, 2), np.int32),
                                                     {"scale": (W / 1000.0, H / 1000.0), "is_sparse": False}))
    rec0 = copy.deepcopy(rec)
    cut = 12
    conf = {"strategy": "costmaps", "costmaps": {"dense_cut_size": cut}, "optimizer": {"solver": {"max_num_iterations": 6}}}
    out = ba_pkg.BundleAdjuster.create(conf).refine_multilevel(rec, fm)
    s, cmaps, refs = out["summary"][0], out["costmaps"][0], out["references"][0]
    assert s.final_cost < s.initial_cost and cmaps.channels == 3
    for image in rec0.images.values():
        cm = cmaps.fmap(image.name)
        seen = [k for k, p in enumerate(image.points2D) if p.has_point3D()]
        assert cm.is_sparse and cm.point2D_ids == seen and cm.shape == (cut, cut, 3)
        cam = rec0.cameras[image.camera_id]
        xy = cameras.world_to_image(cam.model_id, cam.params, image.qvec, image.tvec,
                                    np.array([rec0.points3D[image.points2D[k].point3D_id].xyz for k in seen]))
        uv = xy * np.array([W / 1000.0, H / 1000.0]) - 0.5
        want = np.clip(np.trunc(uv - cut / 2.0), 0, [W - cut, H - cut]).astype(np.int32)       # ToCorner, featurepatch.cc:324-336
        assert np.array_equal(cm.corners, want)
        assert ((uv - want >= 1.0) | (want == 0)).all() and ((uv - want <= cut - 2.0) | (want == W - cut)).all()
    # the references are those of the dense maps themselves
    dense_refs = ba.ReferenceExtractor({"iters": 100, "keep_observations": False}, {}).run(
        [0] * (max(rec0.points3D.keys()) + 1), rec0, fm.fset(0))
    for pid, r in dense_refs.items():
        assert refs[pid].source == r.source and np.abs(refs[pid].descriptor - r.descriptor).max() < 1e-12
    # and the windows hold the map's own values
    sliced = ba.slice_dense_maps(fm.fset(0), rec0, set(rec0.points3D.keys()), cut)
    name = next(iter(fm.fset(0).keys()))
    win, c = sliced.fmap(name).patches[3], sliced.fmap(name).corners[3]
    assert np.array_equal(win, fm.fset(0).fmap(name).patches[0][c[1]:c[1] + cut, c[0]:c[0] + cut])


def test_keypoint_adjustment_on_a_lazily_filled_cache(oracle_engine, tmp_path):
    """KA on FeatureManager(path, fill=False): the maps a solve touches are resident for its duration only, and the refined
    keypoints are the ones the in-memory manager gives."""
    import copy
    from pixsfm import base, features, keypoint_adjustment as ka_pkg
    from pixsfm.features import store_features
    from pixsfm.util import synthetic
    sc = synthetic.make_ka_scene(n_images=5, n_tracks=30, track_len=3, channels=16, seed=5, kp_sigma=1.0)
    g = base.Graph()
    names = ["im%d" % i for i in range(5)]
    keypoints = {names[i]: np.ascontiguousarray(sc["keypoints"][sc["node_image"] == i]) for i in range(5)}
    for n in range(len(sc["node_image"])):
        g.add_node(names[sc["node_image"][n]], int(sc["node_feature"][n]))
    for e in range(len(sc["edge_src"])):
        g.add_edge(g.nodes[sc["edge_src"][e]], g.nodes[sc["edge_dst"][e]], sc["edge_sim"][e])
    fm = features.FeatureManager([16], np.float16)
    for i in range(5):
        m = np.where(sc["node_image"] == i)[0]
        fm.fset(0).emplace(names[i], features.FeatureMap(np.ascontiguousarray(sc["patches"][m]), sc["node_feature"][m].tolist(),
                                                          sc["corner"][m], {"scale": sc["scale"][m[0]], "is_sparse": True}))
    store_features.write_feature_manager_cache(tmp_path / "ka.h5", fm)
    lazy = store_features.load_features_from_cache(tmp_path / "ka.h5", fill=False)
    kp_a, kp_b = copy.deepcopy(keypoints), copy.deepcopy(keypoints)
    conf = {"max_kps_per_problem": 20}
    ka_pkg.KeypointAdjuster.create(conf).refine_multilevel(kp_a, fm, g)
    ka_pkg.KeypointAdjuster.create(conf).refine_multilevel(kp_b, lazy, g)
    assert any(np.abs(kp_a[n] - keypoints[n]).max() > 1e-3 for n in names)
    for name in names:
        assert np.array_equal(kp_a[name], kp_b[name])
    assert not any(m.is_loaded for m in lazy.fset(0)._maps.values())


def test_query_localizer_reference_surface(oracle_engine):
    """QueryLocalizer with the reference's constructor options and `localize` signature (localization/main.py:301-497):
    references extracted at construction from a FeatureManager, the four target_reference kinds, unique inliers by minimal
    reprojection error, inliers recomputed from the final pose."""
    import copy
    import test_gpu_localization as T
    from pixsfm import localization as loc_pkg
    from pixsfm.localization import main as M
    from pixsfm.util import cameras
    rec, fm, refs, qid, fmap, p2D_idxs, p3D_ids, kps = T._scene()
    img = rec.images[qid]
    cam = copy.deepcopy(rec.cameras[img.camera_id])
    pose = {"success": True, "qvec": img.qvec.copy(), "tvec": img.tvec.copy(), "inliers": [True] * (len(kps) + 2)}
    pnp = lambda p2, p3, c: dict(pose, qvec=pose["qvec"].copy(), tvec=pose["tvec"].copy())        # noqa: E731
    ql = loc_pkg.QueryLocalizer(rec, {"max_tracks_per_problem": 10}, dense_features=fm, pose_estimator=pnp)
    assert len(ql.references) == 1 and sorted(ql.references[0]) == sorted(refs)
    for pid in list(refs)[:5]:                                  # the same references as extracted by hand
        assert ql.references[0][pid].source == refs[pid].source
        assert np.abs(ql.references[0][pid].descriptor - refs[pid].descriptor).max() < 1e-12
    for kind, check in (("nearest", lambda r: isinstance(r, np.ndarray)), ("robust_mean", lambda r: isinstance(r, np.ndarray)),
                        ("all_observations", lambda r: isinstance(r, list) and len(r) == 4),
                        ("full", lambda r: hasattr(r, "descriptor"))):
        q = loc_pkg.QueryLocalizer(rec, {"target_reference": kind}, references=[refs], pose_estimator=pnp)
        got = q.get_query_references(p3D_ids, [fmap], kps, p2D_idxs)
        assert len(got) == 1 and len(got[0]) == len(kps) and all(check(r) for r in got[0])
    # two correspondences of ONE keypoint (and two of one 3D point): the one with the smaller reprojection error survives
    all_kps = np.zeros((max(p2D_idxs) + 1, 2)); all_kps[list(p2D_idxs)] = kps
    idxs = list(p2D_idxs) + [p2D_idxs[0], p2D_idxs[1]]
    pids = list(p3D_ids) + [p3D_ids[5], p3D_ids[1]]
    errors = M.compute_reprojection_errors(all_kps[idxs], [rec.points3D[p] for p in pids], img.qvec, img.tvec, cam)
    proj = cameras.world_to_image(cam.model_id, cam.params, img.qvec, img.tvec, np.array([rec.points3D[p].xyz for p in pids]))
    assert np.allclose(errors, np.linalg.norm(proj - all_kps[idxs], axis=1), atol=1e-12)
    keep = M.find_unique_min_reproj_inliers(pids, img.qvec, img.tvec, cam, all_kps[idxs], rec, pre_inliers=[True] * len(idxs),
                                            point2D_idxs=idxs)
    assert keep[0] and not keep[-2]                              # keypoint p2D_idxs[0]: its true point beats point p3D_ids[5]
    assert sum(keep[k] for k in (1, len(idxs) - 1)) == 1         # the duplicate of correspondence 1: exactly one is kept
    out = ql.localize(all_kps, idxs, pids, cam, query_fmaps=[fmap])
    assert out["success"] and len(out["inliers"]) == len(idxs) and out["num_inliers"] == sum(out["inliers"])
    assert not out["inliers"][-2]                                # 12 px threshold on the final pose: the wrong pairing is out
    assert loc_pkg.QueryLocalizer(rec, {}, references=[refs], pose_estimator=pnp).localize(all_kps, [], [], cam, query_fmaps=[fmap]) == {"success": False}
    with pytest.raises(ValueError, match="image_path or query_fmaps"):
        ql.localize(all_kps, idxs, pids, cam)
