"""End-to-end through the reference-facing Python surface (BundleAdjuster / KeypointAdjuster, the names and call
shapes of the reference's pixsfm package) on the GPU, checked against the oracle run on the same problem IR."""
import copy

import numpy as np
import pytest

import oracle_lib as O
from pixsfm import base, bundle_adjustment as ba_pkg, features, keypoint_adjustment as ka_pkg
from pixsfm._pixsfm import _bundle_adjustment as ba
from pixsfm._pixsfm import _capi
from pixsfm.util import synthetic
from recon_util import make_reconstruction

pytestmark = pytest.mark.gpu


def test_bundle_adjuster_refine_multilevel_matches_oracle():
    rec, fm, _, gt = make_reconstruction(n_cams=6, n_points=60, track_len=4, channels=128, seed=21)
    rec_ref = copy.deepcopy(rec)
    conf = {"optimizer": {"solver": {"max_num_iterations": 12}}}
    out = ba_pkg.BundleAdjuster.create(conf).refine_multilevel(rec, fm)
    assert len(out["summary"]) == 1 and len(out["references"]) == 1
    summary, references = out["summary"][0], out["references"][0]
    assert summary.final_cost < summary.initial_cost
    # oracle on the IR the mirror builds from the untouched copy
    setup = ba_pkg.default_problem_setup(rec_ref)
    options = ba.BundleOptimizerOptions()
    fview = features.FeatureView(fm.fset(0), rec_ref)
    ic = _capi.default_interp()
    prob_r, ir_r = ba.build_problem(rec_ref, fview, None, None, None, for_references=set(rec_ref.points3D.keys()))
    refs_o, src_o = O.refs_compute(prob_r, ic, iters=100)
    for k, pid in enumerate(ir_r.point_ids):   # reference source + descriptor: index bit-exact, values tight
        assert references[pid].source == ir_r.obs[int(src_o[k])][:2]
        assert np.abs(references[pid].descriptor.reshape(-1) - refs_o[k]).max() < 1e-12
    oracle_refs = {pid: features.Reference(ir_r.obs[int(src_o[k])][:2], refs_o[k].reshape(1, -1)) for k, pid in enumerate(ir_r.point_ids)}
    prob, ir = ba.build_problem(rec_ref, fview, setup, options, oracle_refs)
    so = _capi.default_ba_options(use_inner_iterations=1, max_num_iterations=12)
    s = O.ba_solve(prob, ic, so)
    assert abs(summary.final_cost - s["final_cost"]) <= 1e-6 * s["final_cost"]
    ba.write_back(rec_ref, prob, ir)
    for i in rec.images:
        assert np.abs(rec.images[i].qvec - rec_ref.images[i].qvec).max() < 1e-6
        assert np.abs(rec.images[i].tvec - rec_ref.images[i].tvec).max() < 1e-6
    for c in rec.cameras:
        assert np.abs(rec.cameras[c].params[1:] - rec_ref.cameras[c].params[1:]).max() < 1e-6
        assert abs(rec.cameras[c].params[0] / rec_ref.cameras[c].params[0] - 1) < 1e-6
    for p in rec.points3D:
        assert np.abs(rec.points3D[p].xyz - rec_ref.points3D[p].xyz).max() < 1e-6
    # gauge: first image untouched, x translation of the second fixed (default_problem_setup)
    assert np.array_equal(rec.images[1].qvec / np.linalg.norm(rec.images[1].qvec), rec.images[1].qvec)


def test_keypoint_adjuster_refine_multilevel_moves_keypoints_towards_truth():
    sc = synthetic.make_ka_scene(n_images=6, n_tracks=50, track_len=4, channels=128, seed=9, kp_sigma=1.0)
    g = base.Graph()
    names = ["im%d" % i for i in range(6)]
    keypoints = {}
    for i in range(6):
        m = sc["node_image"] == i
        keypoints[names[i]] = np.ascontiguousarray(sc["keypoints"][m])
    for n in range(len(sc["node_image"])):   # node n = (image, feature); same insertion order as the flat arrays
        g.add_node(names[sc["node_image"][n]], int(sc["node_feature"][n]))
    for e in range(len(sc["edge_src"])):
        g.add_edge(g.nodes[sc["edge_src"][e]], g.nodes[sc["edge_dst"][e]], sc["edge_sim"][e])
    fm = features.FeatureManager([128], np.float16)
    for i in range(6):
        m = np.where(sc["node_image"] == i)[0]
        fm.fset(0).emplace(names[i], features.FeatureMap(np.ascontiguousarray(sc["patches"][m]),
                                                          sc["node_feature"][m].tolist(), sc["corner"][m],
                                                          {"scale": sc["scale"][m[0]], "is_sparse": True}))
    before = {k: v.copy() for k, v in keypoints.items()}
    out = ka_pkg.KeypointAdjuster.create({"max_kps_per_problem": 20}).refine_multilevel(keypoints, fm, g)
    s = out["summary"][0]
    assert s.final_cost < 0.5 * s.initial_cost
    moved = sum(np.abs(keypoints[k] - before[k]).max() > 1e-3 for k in keypoints)
    assert moved == 6
    # cross-check with the flat-IR path that tests/test_gpu_ka.py pins against the oracle
    from ka_util import make_ka_problem
    from pixsfm._pixsfm import _engine
    prob, _, lab = make_ka_problem(n_images=6, n_tracks=50, track_len=4, channels=128, seed=9, kp_sigma=1.0,
                                   max_per_problem=20, bound=4.0)
    _engine.ka_run(prob, _capi.default_interp(), _capi.default_ka_options())
    flat = np.zeros_like(prob.keypoints)
    for n in range(len(sc["node_image"])):
        flat[n] = keypoints[names[sc["node_image"][n]]][sc["node_feature"][n]]
    assert np.abs(flat - prob.keypoints).max() < 1e-9


def test_device_resident_feature_maps_give_identical_results():
    """FeatureMaps whose patches already live on the GPU (anything with
    __cuda_array_interface__, e.g. the torch tensors the CNN produces) are consumed without a host round trip and
    give the same results as the numpy path (up to the run-to-run order of the fp64 atomics in the block build)."""
    torch = pytest.importorskip("torch")
    rec, fm, _, gt = make_reconstruction(n_cams=6, n_points=50, track_len=4, channels=128, seed=31)
    rec_dev = copy.deepcopy(rec)
    keep = []   # the device tensors must outlive the FeatureMaps (they hold only the pointer's owner)
    fm_dev = features.FeatureManager([128], np.float16)
    for name in fm.fset(0).keys():
        m = fm.fset(0).fmap(name)
        t = torch.from_numpy(m.patches).cuda().contiguous()
        keep.append(t)
        fm_dev.fset(0).emplace(name, features.FeatureMap(t, m.point2D_ids, m.corners, {"scale": m.scale, "is_sparse": True}))
        assert fm_dev.fset(0).fmap(name).patches.ptr == t.data_ptr()
    conf = {"optimizer": {"solver": {"max_num_iterations": 8}}}
    out_h = ba_pkg.BundleAdjuster.create(conf).refine_multilevel(rec, fm)
    out_d = ba_pkg.BundleAdjuster.create(conf).refine_multilevel(rec_dev, fm_dev)
    ch, cd = out_h["summary"][0].final_cost, out_d["summary"][0].final_cost
    assert abs(cd - ch) <= 1e-10 * ch, (cd, ch)
    # the host path brings over one 8x8 window per observation (window residency), the device path only the problem tables
    assert out_d["summary"][0].h2d_bytes < 0.05 * out_h["summary"][0].h2d_bytes, (out_d["summary"][0].h2d_bytes, out_h["summary"][0].h2d_bytes)
    for p in rec.points3D:
        assert np.abs(rec.points3D[p].xyz - rec_dev.points3D[p].xyz).max() < 1e-10
    # cost-map strategy on the device-resident maps too
    out_c = ba_pkg.BundleAdjuster.create({"strategy": "costmaps", **conf}).refine_multilevel(copy.deepcopy(rec_dev), fm_dev)
    assert out_c["summary"][0].final_cost < out_c["summary"][0].initial_cost


def test_topological_reference_keypoint_adjuster_matches_flat_oracle():
    """strategy 'topological_reference' (keypoint_adjustment/main.py:206-250): root edges only, unit weights, root
    regularisation — the edge list the mirror enumerates, solved on the GPU, equals the oracle on the same IR."""
    from pixsfm._pixsfm import _keypoint_adjustment as ka
    sc = synthetic.make_ka_scene(n_images=6, n_tracks=40, track_len=4, channels=128, seed=12, kp_sigma=1.0)
    g = base.Graph()
    names = ["im%d" % i for i in range(6)]
    keypoints = {names[i]: np.ascontiguousarray(sc["keypoints"][sc["node_image"] == i]) for i in range(6)}
    for n in range(len(sc["node_image"])):
        g.add_node(names[sc["node_image"][n]], int(sc["node_feature"][n]))
    for e in range(len(sc["edge_src"])):
        g.add_edge(g.nodes[sc["edge_src"][e]], g.nodes[sc["edge_dst"][e]], sc["edge_sim"][e])
    fm = features.FeatureManager([128], np.float16)
    for i in range(6):
        m = np.where(sc["node_image"] == i)[0]
        fm.fset(0).emplace(names[i], features.FeatureMap(np.ascontiguousarray(sc["patches"][m]), sc["node_feature"][m].tolist(),
                                                          sc["corner"][m], {"scale": sc["scale"][m[0]], "is_sparse": True}))
    track_labels = base.compute_track_labels(g)
    root_labels = base.compute_root_labels(g, track_labels, base.compute_score_labels(g, track_labels))
    before = {k: v.copy() for k, v in keypoints.items()}
    kp_fm = {k: v.copy() for k, v in keypoints.items()}
    out = ka_pkg.KeypointAdjuster.create({"strategy": "topological_reference", "max_kps_per_problem": 20}).refine_multilevel(
        keypoints, fm, g, track_labels, root_labels)
    out_fm = ka_pkg.KeypointAdjuster.create({"max_kps_per_problem": 20}).refine_multilevel(kp_fm, fm, g, track_labels, root_labels)
    s, s_fm = out["summary"][0], out_fm["summary"][0]
    assert s.final_cost < s.initial_cost
    # linear instead of quadratic number of residuals: every non-root node has exactly one edge, to its root
    n_roots = int(np.sum(root_labels))
    assert s.num_residual_blocks == len(g.nodes) - n_roots
    assert s.num_residual_blocks < s_fm.num_residual_blocks
    # roots are constant, everything else moved
    for n, node in enumerate(g.nodes):
        name = g.image_id_to_name[node.image_id]
        d = np.abs(keypoints[name][node.feature_idx] - before[name][node.feature_idx]).max()
        assert (d == 0.0) if root_labels[n] else (d > 0.0)


def test_dense_feature_maps_match_oracle():
    """Dense FeatureMaps (one kDenseId patch per image answering for every keypoint, featuremap.h:104-119) run through
    the same kernels: many observations share one large patch.  GPU vs oracle on the IR the mirror builds."""
    rec, fm_sparse, prob0, gt = make_reconstruction(n_cams=5, n_points=40, track_len=3, channels=16, seed=41)
    rng = np.random.default_rng(7)
    H = W = 72
    fm = features.FeatureManager([16], np.float16)
    for name in fm_sparse.fset(0).keys():
        # a smooth random field so that the bicubic gradients are informative
        base_f = rng.normal(size=(H // 8 + 2, W // 8 + 2, 16))
        yy, xx = np.meshgrid(np.linspace(1, H // 8, H), np.linspace(1, W // 8, W), indexing="ij")
        y0, x0 = yy.astype(int), xx.astype(int)
        fy, fx = (yy - y0)[..., None], (xx - x0)[..., None]
        dense = ((1 - fy) * (1 - fx) * base_f[y0, x0] + (1 - fy) * fx * base_f[y0, x0 + 1]
                 + fy * (1 - fx) * base_f[y0 + 1, x0] + fy * fx * base_f[y0 + 1, x0 + 1]).astype(np.float16)
        fm.fset(0).emplace(name, features.FeatureMap(np.ascontiguousarray(dense[None]), [features.kDenseId], np.zeros((1, 2), np.int32),
                                                     {"scale": (W / 1000.0, H / 1000.0), "is_sparse": False}))
    rec_ref = copy.deepcopy(rec)
    conf = {"optimizer": {"solver": {"max_num_iterations": 8}}}
    out = ba_pkg.BundleAdjuster.create(conf).refine_multilevel(rec, fm)
    s = out["summary"][0]
    fview = features.FeatureView(fm.fset(0), rec_ref)
    ic = _capi.default_interp()
    prob_r, ir_r = ba.build_problem(rec_ref, fview, None, None, None, for_references=set(rec_ref.points3D.keys()))
    assert prob_r.n_patches == 5 and prob_r.n_obs == 120 and (prob_r.ph, prob_r.pw) == (H, W)
    refs_o, src_o = O.refs_compute(prob_r, ic, iters=100)
    oracle_refs = {pid: features.Reference(ir_r.obs[int(src_o[k])][:2], refs_o[k].reshape(1, -1)) for k, pid in enumerate(ir_r.point_ids)}
    prob, ir = ba.build_problem(rec_ref, fview, ba_pkg.default_problem_setup(rec_ref), ba.BundleOptimizerOptions(), oracle_refs)
    so = _capi.default_ba_options(use_inner_iterations=1, max_num_iterations=8)
    s_o = O.ba_solve(prob, ic, so)
    assert s.num_iterations == s_o["num_iterations"]
    assert abs(s.initial_cost - s_o["initial_cost"]) <= 1e-10 * s_o["initial_cost"]
    # random (unmatched) feature fields make a badly conditioned problem: rounding differences grow along the LM path
    assert abs(s.final_cost - s_o["final_cost"]) <= 1e-4 * s_o["final_cost"]


def test_keypoint_adjustment_from_a_colmap_database(tmp_path):
    """refine_colmap.py:105-112 shape: keypoints + matches come out of a COLMAP database, the matching graph is built
    from them, the keypoints are adjusted on the GPU and written back."""
    from pixsfm.util import colmap as cio
    from pixsfm.util.database import COLMAPDatabase
    sc = synthetic.make_ka_scene(n_images=6, n_tracks=50, track_len=4, channels=128, seed=15, kp_sigma=1.0)
    names = ["im%d.jpg" % i for i in range(6)]
    path = tmp_path / "database.db"
    db = COLMAPDatabase.connect(path)
    db.create_tables()
    cam = db.add_camera(2, 1000, 1000, [1200.0, 500.0, 500.0, 0.0])
    ids = [db.add_image(n, cam) for n in names]
    for i in range(6):
        db.add_keypoints(ids[i], sc["keypoints"][sc["node_image"] == i])
    by_pair = {}
    for s, d in zip(sc["edge_src"], sc["edge_dst"]):
        a, b = int(sc["node_image"][s]), int(sc["node_image"][d])
        fa, fb = int(sc["node_feature"][s]), int(sc["node_feature"][d])
        if a > b:
            a, b, fa, fb = b, a, fb, fa
        by_pair.setdefault((a, b), []).append((fa, fb))
    for (a, b), m in by_pair.items():
        db.add_matches(ids[a], ids[b], np.array(m, np.uint32))
    db.commit(); db.close()

    keypoints = cio.read_keypoints_from_db(path)
    pairs, matches, scores = cio.read_matches_from_db(path)
    assert scores is None and sum(len(m) for m in matches) == len(sc["edge_src"])
    graph = ka_pkg.build_matching_graph(pairs, matches, scores)
    needed = ka_pkg.extract_patchdata_from_graph(graph)
    fm = features.FeatureManager([128], np.float16)
    for i in range(6):
        m = np.where(sc["node_image"] == i)[0]
        assert set(needed[names[i]]) <= set(sc["node_feature"][m].tolist())
        fm.fset(0).emplace(names[i], features.FeatureMap(np.ascontiguousarray(sc["patches"][m]), sc["node_feature"][m].tolist(),
                                                          sc["corner"][m], {"scale": sc["scale"][m[0]], "is_sparse": True}))
    before = {k: v.copy() for k, v in keypoints.items()}
    out = ka_pkg.KeypointAdjuster.create({"max_kps_per_problem": 20}).refine_multilevel(keypoints, fm, graph)
    s = out["summary"][0]
    assert s.final_cost < 0.5 * s.initial_cost
    assert all(np.abs(keypoints[k] - before[k]).max() > 1e-3 for k in keypoints)
    cio.write_keypoints_to_db(path, keypoints)
    back = cio.read_keypoints_from_db(path)
    for k in keypoints:
        assert np.array_equal(back[k], keypoints[k].astype(np.float32).astype(np.float64))
    # the same through the PixSfM driver, into a second database (refine_colmap.py:99-115)
    from pixsfm.refine_colmap import PixSfM
    db = COLMAPDatabase.connect(path)
    db.execute("DELETE FROM keypoints")
    for i in range(6):
        db.add_keypoints(ids[i], sc["keypoints"][sc["node_image"] == i])
    db.commit(); db.close()
    out_db = tmp_path / "refined.db"
    kp2, ka_data, _ = PixSfM({"KA": {"max_kps_per_problem": 20}}).refine_keypoints_from_db(out_db, path, feature_manager=fm)
    assert ka_data["summary"][0].final_cost < 0.5 * ka_data["summary"][0].initial_cost
    refined = cio.read_keypoints_from_db(out_db)
    untouched = cio.read_keypoints_from_db(path)
    for k in kp2:
        assert np.abs(kp2[k] - keypoints[k]).max() < 1e-9                   # same solve as the direct call above
        assert np.array_equal(refined[k], kp2[k].astype(np.float32).astype(np.float64))
        assert np.array_equal(untouched[k], before[k])                      # the input database is left alone
    assert cio.read_matches_from_db(out_db)[0] == pairs


def test_pixsfm_refines_a_colmap_model_directory(tmp_path):
    """PixSfM.refine_reconstruction (refine_colmap.py:117-131): model files -> BA on the GPU -> model files, equal to
    running the adjuster on the in-memory reconstruction."""
    from pixsfm.refine_colmap import PixSfM
    from pixsfm.util.colmap_types import Reconstruction
    rec, fm, _, gt = make_reconstruction(n_cams=6, n_points=60, track_len=4, channels=128, seed=23)
    rec.write(str(tmp_path / "in"))
    conf = {"BA": {"optimizer": {"solver": {"max_num_iterations": 8}}}}
    sfm = PixSfM(conf)
    out_rec, ba_data, fm_out = sfm.refine_reconstruction(tmp_path / "out", tmp_path / "in", feature_manager=fm)
    assert fm_out is fm and ba_data["summary"][0].final_cost < ba_data["summary"][0].initial_cost
    direct = copy.deepcopy(rec)
    ba_pkg.BundleAdjuster.create(conf["BA"]).refine_multilevel(direct, fm)
    back = Reconstruction.read(tmp_path / "out")
    for iid in rec.images:
        assert np.array_equal(back.images[iid].qvec, out_rec.images[iid].qvec)          # what was written is what was solved
        assert np.abs(back.images[iid].qvec - direct.images[iid].qvec).max() < 1e-9     # and equals the direct call
        assert np.abs(back.images[iid].tvec - direct.images[iid].tvec).max() < 1e-9
    for pid in rec.points3D:
        assert np.abs(back.points3D[pid].xyz - direct.points3D[pid].xyz).max() < 1e-9
    moved = max(np.abs(back.points3D[p].xyz - rec.points3D[p].xyz).max() for p in rec.points3D)
    assert moved > 1e-6


def test_solver_callbacks_see_every_iteration_and_can_stop_the_solve():
    """ceres::IterationCallback through `solver.callbacks` (reference util/misc.py:30-36): one call per iteration record,
    starting with iteration 0; the trajectory equals the one-shot solve's; a callback can terminate the solve."""
    rec, fm, _, _ = make_reconstruction(n_cams=6, n_points=60, track_len=4, channels=128, seed=23)
    rec_a, rec_b, rec_c = copy.deepcopy(rec), copy.deepcopy(rec), copy.deepcopy(rec)
    conf = {"optimizer": {"solver": {"max_num_iterations": 8}}}
    plain = ba_pkg.BundleAdjuster.create(conf).refine_multilevel(rec_a, fm)["summary"][0]

    seen = []
    adj = ba_pkg.BundleAdjuster.create(conf)
    adj.callbacks = [lambda it: seen.append((it["iteration"], it["cost"]))]
    stepped = adj.refine_multilevel(rec_b, fm)["summary"][0]
    assert [i for i, _ in seen] == list(range(len(seen))) and len(seen) == len(stepped.iterations)
    assert len(seen) == len(plain.iterations)
    assert abs(stepped.final_cost - plain.final_cost) <= 1e-8 * plain.final_cost
    for p in rec_a.points3D:
        assert np.abs(rec_a.points3D[p].xyz - rec_b.points3D[p].xyz).max() < 1e-7

    calls = []
    def stop_after_two(it):
        calls.append(it["iteration"])
        return 2 if it["iteration"] >= 2 else 0
    adj = ba_pkg.BundleAdjuster.create(conf)
    adj.callbacks = [stop_after_two]
    stopped = adj.refine_multilevel(rec_c, fm)["summary"][0]
    assert calls == [0, 1, 2]
    assert "SOLVER_TERMINATE_SUCCESSFULLY" in stopped.message
    assert stopped.final_cost < stopped.initial_cost and stopped.final_cost >= plain.final_cost
