"""CPU tests of the reference-facing Python mirror: the problem construction / parameterisation semantics of
BundleOptimizer<D> (the cases of the reference's bundle_optimizer_test.cc:163-353, which pins them with the
geometric residual) resolved into the constancy masks of the problem IR; configuration plumbing and error
behaviour.  No GPU needed: nothing is solved here."""
import numpy as np
import pytest

from pixsfm import base, bundle_adjustment as ba_pkg, keypoint_adjustment as ka_pkg
from pixsfm._pixsfm import _bundle_adjustment as ba
from pixsfm._pixsfm import _capi
from pixsfm._pixsfm._features import FeatureView
from recon_util import make_reconstruction


def _build(rec, fm, setup, **opt):
    options = ba.BundleOptimizerOptions(opt)
    fview = FeatureView(fm.fset(0), rec)
    return ba.build_problem(rec, fview, setup, options, references=None)


def test_default_setup_two_view_gauge():
    rec, fm, prob0, _ = make_reconstruction(n_cams=2, n_points=30, track_len=2, channels=16)
    setup = ba_pkg.default_problem_setup(rec)            # TestTwoView: pose(0) constant, tvec(1).x constant
    prob, ir = _build(rec, fm, setup)
    assert ir.image_ids == [1, 2] and len(ir.obs) == 60
    assert prob.pose_const.tolist() == [1, 0] and prob.tvec_const_mask.tolist() == [0, 1]
    assert prob.point_const.sum() == 0
    assert prob.cam_const_mask.tolist() == [0x6, 0x6]   # SIMPLE_RADIAL: principal point (cx,cy) constant by default
    assert np.all(np.diff(prob.obs_pt) >= 0)
    # observations of one point keep the image enumeration order (ascending image id)
    assert prob.obs_img[:2].tolist() == [0, 1]


def test_two_view_constant_camera_and_refine_flags():
    rec, fm, _, _ = make_reconstruction(n_cams=2, n_points=20, track_len=2, channels=16)
    setup = ba_pkg.default_problem_setup(rec)
    setup.set_constant_camera(1)                         # TestTwoViewConstantCamera
    prob, _ = _build(rec, fm, setup)
    assert prob.cam_const_mask.tolist() == [0xFFFFFFFF, 0x6]
    prob, _ = _build(rec, fm, setup, refine_focal_length=False, refine_extra_params=False)
    assert prob.cam_const_mask.tolist() == [0xFFFFFFFF, 0xFFFFFFFF]
    prob, _ = _build(rec, fm, ba_pkg.default_problem_setup(rec), refine_principal_point=True, refine_extra_params=False)
    assert prob.cam_const_mask.tolist() == [0x8, 0x8]
    prob, _ = _build(rec, fm, ba_pkg.default_problem_setup(rec), refine_extrinsics=False)
    assert prob.pose_const.tolist() == [1, 1]


def test_partially_contained_tracks_become_constant_points():
    rec, fm, _, _ = make_reconstruction(n_cams=3, n_points=25, track_len=3, channels=16)
    setup = ba.BundleAdjustmentSetup()                   # TestPartiallyContainedTracks: only images 1 and 2 in the problem
    setup.add_images({1, 2})
    setup.set_constant_pose(1)
    setup.set_constant_tvec(2, [0])
    prob, ir = _build(rec, fm, setup)
    assert ir.image_ids == [1, 2] and len(ir.obs) == 50
    assert prob.point_const.all()                        # every track has a third element outside the problem
    setup.add_variable_point(5)                          # TestPartiallyContainedTracksForceToOptimizePoint
    prob, ir = _build(rec, fm, setup)
    assert ir.image_ids == [1, 2, 3] and len(ir.obs) == 51
    k = ir.point_ids.index(5)
    assert prob.point_const[k] == 0 and prob.point_const.sum() == 24
    assert prob.pose_const.tolist() == [1, 0, 1]         # image 3 is not in the setup: pose constant
    assert prob.cam_const_mask[ir.camera_ids.index(rec.images[3].camera_id)] == 0xFFFFFFFF
    setup2 = ba.BundleAdjustmentSetup(); setup2.add_images({1, 2, 3}); setup2.add_constant_point(7)   # TestConstantPoints
    prob, ir = _build(rec, fm, setup2)
    assert prob.point_const.sum() == 1 and prob.point_const[ir.point_ids.index(7)] == 1


def test_min_track_length_and_empty_problem():
    rec, fm, _, _ = make_reconstruction(n_cams=3, n_points=10, track_len=2, channels=16)
    setup = ba.BundleAdjustmentSetup(); setup.add_images({1, 2, 3})
    prob, ir = _build(rec, fm, setup, min_track_length=3)
    assert prob.n_obs == 0                               # all tracks have length 2 -> skipped (bundle_optimizer.h:262-265)
    opt = ba.FeatureReferenceBundleOptimizer({"min_track_length": 3}, setup, base.interpolation_default_conf)
    assert opt.run(rec, FeatureView(fm.fset(0), rec), {}) is False   # returns False when there are no residuals


def test_setup_error_behaviour_matches_reference():
    s = ba.BundleAdjustmentSetup(); s.add_image(1)
    with pytest.raises(ValueError):
        s.set_constant_pose(2)                           # image not in setup
    s.set_constant_pose(1)
    with pytest.raises(ValueError):
        s.set_constant_tvec(1, [0])                      # already constant pose
    s.add_image(2)
    with pytest.raises(ValueError):
        s.set_constant_tvec(2, [0, 0])                   # duplicate indices
    s.add_variable_point(3)
    with pytest.raises(ValueError):
        s.add_constant_point(3)
    with pytest.raises(ValueError):
        ba.BundleOptimizerOptions({"no_such_option": 1})    # strict keys (helpers.h:149-232)
    with pytest.raises(ValueError):
        base.InterpolationConfig({"nodes": [[0, 0], [1, 1]]}).validate_for_device()
    with pytest.raises(ValueError):
        ba_pkg.BundleAdjuster.create({"strategy": "patch_warp"})
    opt = ba.FeatureReferenceBundleOptimizer({}, s, base.interpolation_default_conf)
    with pytest.raises(ValueError):
        opt.run(None, None, {})


def test_python_defaults_are_the_reference_defaults():
    conf = ba_pkg.BundleAdjuster.create({}).conf
    assert conf.optimizer.solver.use_inner_iterations is True and conf.references.iters == 100
    assert conf.max_tracks_per_problem == 10 and conf.optimizer.loss.params == [0.25]
    assert conf.optimizer.refine_principal_point is False and conf.interpolation.l2_normalize is True
    conf = ka_pkg.KeypointAdjuster.create({"optimizer": {"bound": 2.0}}).conf
    assert conf.optimizer.bound == 2.0 and conf.optimizer.solver.parameter_tolerance == 1.0e-5
    assert conf.max_kps_per_problem == 50 and conf.optimizer.weight_by_sim is True
    rec, fm, _, _ = make_reconstruction(n_cams=2, n_points=25, track_len=2, channels=16)
    labels = ba_pkg.find_problem_labels(rec, 10)
    assert labels[0] == -1 and labels[1] == 0 and labels[10] == 1 and labels[25] == 2


def test_graph_mirror_and_labels():
    g = base.Graph()
    g.register_matches("a", "b", np.array([[0, 0], [1, 1], [2, 2]]), np.array([0.9, 0.8, 0.7]))
    g.register_matches("b", "c", np.array([[0, 0], [1, 2]]), np.array([0.6, 0.95]))
    g.register_matches("a", "c", np.array([[0, 1]]), np.array([0.5]))   # would put two features of c into one track
    assert g.image_name_to_id == {"a": 0, "b": 1, "c": 2} and len(g.nodes) == 9
    tl = base.compute_track_labels(g)
    sc = base.compute_score_labels(g, tl)
    rt = base.compute_root_labels(g, tl, sc)
    node = {(n.image_id, n.feature_idx): n.node_idx for n in g.nodes}
    assert tl[node[(0, 0)]] == tl[node[(1, 0)]] == tl[node[(2, 0)]]
    assert tl[node[(2, 1)]] != tl[node[(0, 0)]]          # refused: track already has a feature in image c
    assert tl[node[(0, 1)]] == tl[node[(1, 1)]] == tl[node[(2, 2)]]
    assert sum(rt) == len(set(tl))
    labels, bins = ka_pkg.find_problem_labels(tl, 50)
    assert len(bins) == 1 and set(labels) == {0}


def test_costmap_strategy_surface_and_defaults():
    """bundle_adjustment/main.py:218-238 + costmap_extractor.h:18-40: same class names, keys and defaults."""
    from pixsfm import bundle_adjustment as ba_pkg
    from pixsfm._pixsfm import _bundle_adjustment as ba
    adj = ba_pkg.BundleAdjuster.create({"strategy": "costmaps"})
    assert isinstance(adj, ba_pkg.CostMapBundleAdjuster)
    assert adj.conf.costmaps.loss.name == "trivial" and adj.conf.costmaps.as_gradientfield is True
    assert adj.conf.costmaps.compute_cross_derivative is False
    cfg = ba.CostMapConfig()
    assert (cfg.upsampling_factor, cfg.as_gradientfield, cfg.compute_cross_derivative, cfg.apply_sqrt, cfg.dense_cut_size) == \
        (1.0, True, False, False, 12)
    assert cfg.get_effective_channels() == 3
    assert ba.CostMapConfig({"as_gradientfield": False}).get_effective_channels() == 1
    assert ba.CostMapConfig({"compute_cross_derivative": True}).get_effective_channels() == 4
    with pytest.raises(ValueError):
        ba.CostMapConfig({"no_such_key": 1})
    with pytest.raises(ValueError):
        ba_pkg.BundleAdjuster.create({"strategy": "patch_warp"})
    c = _capi.default_costmap_config()
    assert (c.loss_type, c.as_gradientfield, c.apply_sqrt, c.upsampling_factor, c.ref_loss_type, c.ref_loss_scale, c.ref_iters) == \
        (0, 1, 0, 1.0, 1, 0.25, 100)


def test_topological_reference_ka_surface():
    """keypoint_adjustment/main.py:206-250 + topological_reference_keypoint_optimizer.h:9-16"""
    from pixsfm._pixsfm import _keypoint_adjustment as ka
    adj = ka_pkg.KeypointAdjuster.create({"strategy": "topological_reference"})
    assert isinstance(adj, ka_pkg.TopologicalReferenceKeypointAdjuster)
    o = ka.TopologicalReferenceKeypointOptimizerOptions()
    assert (o.weight_by_sim, o.root_regularize_weight, o.root_edges_only) == (False, 1.0, True)
    o2 = ka.KeypointOptimizerOptions()
    assert (o2.weight_by_sim, o2.root_regularize_weight, o2.root_edges_only) == (True, -1.0, False)
    opt = ka.TopologicalReferenceKeypointOptimizer({"bound": 2.0}, ka.KeypointAdjustmentSetup(), {})
    assert opt.options.root_edges_only and opt.options.bound == 2.0


def test_localization_surface_and_defaults():
    """localization/main.py:92-262 + query_refinement_options.h: class names, config keys, defaults."""
    from pixsfm import localization as L
    qka, qba = L.QueryKeypointAdjuster(), L.QueryBundleAdjuster()
    assert qka.conf.optimizer.bound == 4.0 and qka.conf.optimizer.loss.name == "trivial"
    assert qka.conf.optimizer.solver.parameter_tolerance == 1e-05 and qka.conf.stack_correspondences is False
    assert qba.conf.optimizer.loss.name == "cauchy" and qba.conf.optimizer.refine_focal_length is False
    o = L.QueryKeypointOptimizerOptions()
    assert o.bound == -1.0 and o.solver["parameter_tolerance"] == 1.0e-4 and o.loss["name"] == "trivial"
    b = L.QueryBundleOptimizerOptions()
    assert b.solver["parameter_tolerance"] == 1.0e-5 and (b.refine_focal_length, b.refine_principal_point, b.refine_extra_params) == (False,) * 3
    assert L.find_unique_inliers([3, 3, 5, 3, 5], pre_inliers=[False, True, True, True, True]) == [False, True, True, False, False]
    assert L.find_unique_min_by_group([0.5, 0.2, 0.9, 0.1], [7, 7, 8, 8]) == [False, True, False, True]
    with pytest.raises(ValueError):
        L.QueryLocalizer(None, {}, references=None)
    with pytest.raises(ValueError):
        qka.solver.run(np.zeros((2, 2)), None, [None])          # references.size() != keypoints.rows()


def test_count_edges_AB_and_track_edges():
    from pixsfm import base
    g = base.Graph()
    for im, f in [("a", 0), ("b", 0), ("c", 0), ("a", 1), ("b", 1)]:
        g.add_node(im, f)
    for s, d in [(0, 1), (1, 2), (0, 2), (3, 4), (2, 3)]:
        g.add_edge(g.nodes[s], g.nodes[d], 1.0)
    labels, roots = [0, 0, 0, 1, 1], [True, False, False, False, True]
    assert base.count_track_edges(g, labels) == [3, 1]
    ab = base.count_edges_AB(g, labels, roots)
    assert len(ab) == 5 and ab[0] == (2, 1) and ab[1] == (1, 0) and ab[2:] == [(0, 0)] * 3      # the 2-3 edge is inter-track
