"""GPU parity for the cost-map path: CostMapExtractor (costmap_extractor.h:230-358) and
CostMapBundleOptimizer (costmap_bundle_optimizer.h:76-132), through the C-ABI, against the CPU oracle."""
import copy

import numpy as np
import pytest

import oracle_lib as O
from pixsfm import bundle_adjustment as ba_pkg, features
from pixsfm._pixsfm import _bundle_adjustment as ba
from pixsfm._pixsfm import _capi, _engine
from pixsfm.util import synthetic
from recon_util import make_reconstruction

pytestmark = pytest.mark.gpu


def _scene(**kw):
    args = dict(n_cams=6, n_points=40, track_len=4, channels=128, seed=1)
    args.update(kw)
    prob, gt = synthetic.make_ba_scene(**args)
    ic = _capi.default_interp()
    refs, _ = O.refs_compute(prob, ic)
    prob.refs = refs
    return prob, ic


def _half_ulps(a, b):
    """distance in fp16 representable steps (monotone integer mapping of the bit patterns)"""
    def key(x):
        u = x.view(np.uint16).astype(np.int32)
        return np.where(u & 0x8000, -(u & 0x7FFF), u)
    return np.abs(key(a) - key(b))


@pytest.mark.parametrize("grad,sqrt,loss", [(1, 0, (0, 1.0)), (1, 1, (1, 0.25)), (0, 0, (0, 1.0)), (0, 1, (2, 0.1))])
def test_costmap_extraction_matches_oracle(grad, sqrt, loss):
    prob, ic = _scene()
    ref = O.costmaps_compute(prob, loss[0], loss[1], bool(grad), bool(sqrt))
    cfg = _capi.default_costmap_config(loss_type=loss[0], loss_scale=loss[1], as_gradientfield=grad, apply_sqrt=sqrt,
                                       compute_refs=0)
    got = _engine.costmaps_compute(prob, ic, cfg, refs=prob.refs)["costmaps"]
    assert got.shape == ref.shape == (prob.n_obs, 16, 16, 3 if grad else 1) and got.dtype == np.float16
    ulps = _half_ulps(got, ref)
    # fp64 dot products in a different summation order, then two roundings to fp16: identical except for
    # the rare value that sits on a rounding boundary
    assert ulps.max() <= 1 and (ulps == 0).mean() > 0.999
    assert np.isfinite(got.astype(np.float32)).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_costmap_extraction_other_dtypes(dtype):
    prob, ic = _scene(channels=16, dtype=dtype, n_points=20)
    ref = O.costmaps_compute(prob)
    got = _engine.costmaps_compute(prob, ic, _capi.default_costmap_config(compute_refs=0), refs=prob.refs)["costmaps"]
    assert got.dtype == dtype
    assert np.allclose(got, ref, rtol=1e-6 if dtype == np.float32 else 1e-12, atol=1e-7 if dtype == np.float32 else 1e-14)


def test_costmaps_with_fused_reference_extraction():
    prob, ic = _scene()
    want_refs, want_src = prob.refs.copy(), O.refs_compute(prob, ic)[1]
    prob.refs = None
    out = _engine.costmaps_compute(prob, ic, _capi.default_costmap_config(), to_device=True)
    assert np.array_equal(out["src_obs"], want_src)
    assert np.abs(out["refs"] - want_refs).max() < 1e-12
    prob.refs = want_refs
    ref = O.costmaps_compute(prob)
    assert _half_ulps(out["costmaps"], ref).max() <= 1
    # the device-resident copy feeds the cost-map BA without a host round trip
    cm_dev = prob.with_patches(out["device_ptr"], on_device=True, patch_shape=out["costmaps"].shape, patch_dtype=0)
    cm_host = prob.with_patches(out["costmaps"])
    ic2 = _capi.default_interp(); ic2.l2_normalize = 0
    so = _capi.default_ba_options(use_inner_iterations=0)
    c_dev = _engine.BAHandle(cm_dev, ic2, so).evaluate()["cost"]
    c_host = _engine.BAHandle(cm_host, ic2, so).evaluate()["cost"]
    assert c_dev == c_host
    _engine.device_free(out["device_ptr"])


@pytest.mark.parametrize("channels", [3, 1])
def test_costmap_ba_blocks_and_step_match_oracle(channels):
    prob, ic = _scene()
    cm = O.costmaps_compute(prob, as_gradientfield=(channels == 3))
    p = prob.with_patches(cm)
    ic2 = _capi.default_interp(); ic2.l2_normalize = 0
    so = _capi.default_ba_options(use_inner_iterations=0)
    ref = O.ba_evaluate(p, ic2, so, residuals=True)
    got = _engine.BAHandle(p, ic2, so).evaluate(residuals=True)
    assert abs(got["cost"] - ref["cost"]) <= 1e-12 * ref["cost"]
    assert np.abs(got["residuals"] - ref["residuals"]).max() <= 1e-12 * np.abs(ref["residuals"]).max()
    lin = O.ba_linearize(p, ic2, so, radius=1e4)
    g = _engine.BAHandle(p, ic2, so).debug_linearize(lin["nc"], lin["nl"], radius=1e4)
    assert np.allclose(g["gc"], lin["gc"], rtol=1e-8, atol=1e-10 * np.abs(lin["gc"]).max())
    assert np.allclose(g["delta"], lin["delta"], rtol=1e-5, atol=1e-8 * np.abs(lin["delta"]).max())


@pytest.mark.parametrize("inner", [0, 1])
def test_costmap_ba_full_solve_matches_oracle(inner):
    prob, ic = _scene(n_points=60)
    p = prob.with_patches(O.costmaps_compute(prob))
    ic2 = _capi.default_interp(); ic2.l2_normalize = 0
    so = _capi.default_ba_options(use_inner_iterations=inner, max_num_iterations=12)
    a, b = p.copy(), p.copy()
    s_ref = O.ba_solve(a, ic2, so); s_gpu = _engine.ba_run(b, ic2, so)
    assert s_gpu["num_iterations"] == s_ref["num_iterations"]
    assert abs(s_gpu["final_cost"] - s_ref["final_cost"]) <= 1e-6 * s_ref["final_cost"]
    assert s_ref["final_cost"] < s_ref["initial_cost"]
    assert np.abs(a.xyz - b.xyz).max() < 1e-6 and np.abs(a.qvec - b.qvec).max() < 1e-6 and np.abs(a.tvec - b.tvec).max() < 1e-6


def test_costmap_bundle_adjuster_strategy_end_to_end():
    rec, fm, _, gt = make_reconstruction(n_cams=6, n_points=60, track_len=4, channels=128, seed=23)
    rec0 = copy.deepcopy(rec)
    conf = {"strategy": "costmaps", "optimizer": {"solver": {"max_num_iterations": 15}}}
    adj = ba_pkg.BundleAdjuster.create(conf)
    assert isinstance(adj, ba_pkg.CostMapBundleAdjuster)
    out = adj.refine_multilevel(rec, fm)
    s, cmaps = out["summary"][0], out["costmaps"][0]
    assert s.final_cost < s.initial_cost
    assert cmaps.channels == 3
    # same problem on the oracle: references -> cost maps -> cost-map BA
    fview = features.FeatureView(fm.fset(0), rec0)
    ic = _capi.default_interp()
    prob_r, ir_r = ba.build_problem(rec0, fview, None, None, None, for_references=set(rec0.points3D.keys()))
    prob_r.refs = O.refs_compute(prob_r, ic, iters=100)[0]
    cm = O.costmaps_compute(prob_r)
    for name, off in ir_r.slab_offsets.items():
        got = cmaps.fmap(name).patches
        used = np.zeros(len(got), bool)
        used[[int(prob_r.obs_patch[k]) - off for k in range(prob_r.n_obs) if off <= prob_r.obs_patch[k] < off + len(got)]] = True
        assert _half_ulps(got[used], cm[off:off + len(got)][used]).max() <= 1
    moved = max(np.abs(rec.points3D[p].xyz - rec0.points3D[p].xyz).max() for p in rec.points3D)
    assert moved > 1e-5
