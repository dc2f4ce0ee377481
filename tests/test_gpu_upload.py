"""The pinned-ring upload of pageable patch memory (csrc/pxr_upload.cu) must deliver exactly the bytes the plain
cudaMemcpy path delivers: single arrays and ragged per-image blocks, chunk sizes that do not divide anything,
repeated calls that reuse the ring."""
import numpy as np
import pytest

from pixsfm._pixsfm import _capi, _engine
from pixsfm.util import synthetic
from ka_util import make_ka_problem

pytestmark = pytest.mark.gpu


@pytest.fixture
def staged(monkeypatch):
    def on(chunk):
        monkeypatch.setenv("PXR_STAGED_UPLOAD_MIN", "1")
        monkeypatch.setenv("PXR_STAGED_CHUNK", str(chunk))

    def off():
        monkeypatch.setenv("PXR_STAGED_UPLOAD_MIN", str(1 << 40))
    return on, off


def _block_problem(prob, sizes):
    """the same problem with its patch slab cut into ragged blocks (what the feature-set mirror hands over)"""
    cuts = np.cumsum([0] + list(sizes))
    assert cuts[-1] == prob.n_patches
    blocks = [np.ascontiguousarray(prob.patches[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    return _capi.BAProblem(prob.cam_model, prob.cam_params, prob.cam_const_mask, prob.qvec, prob.tvec, prob.img_cam,
                           prob.pose_const, prob.tvec_const_mask, prob.xyz, prob.point_const, prob.obs_img, prob.obs_pt,
                           None, prob.corner, prob.scale, refs=prob.refs, obs_patch=prob.obs_patch,
                           patch_blocks=blocks)


@pytest.mark.parametrize("chunk", [4096, 100003, 1 << 20])
def test_staged_upload_single_array_is_exact(staged, chunk):
    on, off = staged
    prob, _ = synthetic.make_ba_scene(n_cams=6, n_points=90, track_len=4, channels=32, seed=11)
    ic = _capi.default_interp()
    off()
    want = _engine.obs_descriptors(prob, ic)
    on(chunk)
    for _ in range(3):      # the ring is reused across calls
        got = _engine.obs_descriptors(prob, ic)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("chunk", [8192, 77777])
def test_staged_upload_ragged_blocks_is_exact(staged, chunk):
    on, off = staged
    prob, _ = synthetic.make_ba_scene(n_cams=5, n_points=80, track_len=4, channels=16, seed=5)
    ic = _capi.default_interp()
    n = prob.n_patches
    sizes = [1, 7, 0, n // 3, 2]
    sizes.append(n - sum(sizes))
    pb = _block_problem(prob, sizes)
    off()
    want = _engine.obs_descriptors(prob, ic)
    assert np.array_equal(_engine.obs_descriptors(pb, ic), want)
    on(chunk)
    assert np.array_equal(_engine.obs_descriptors(pb, ic), want)
    assert np.array_equal(_engine.obs_descriptors(prob, ic), want)


def test_staged_upload_full_solves_match(staged):
    on, off = staged
    prob, _ = synthetic.make_ba_scene(n_cams=6, n_points=60, track_len=4, channels=16, seed=3)
    ic = _capi.default_interp()
    prob.refs = _engine.refs_compute(prob, ic)[0]
    so = _capi.default_ba_options(max_num_iterations=6)
    a, b = prob.copy(), prob.copy()
    off()
    sa = _engine.ba_run(a, ic, so)
    on(50000)
    sb = _engine.ba_run(b, ic, so)
    assert sa["num_iterations"] == sb["num_iterations"]
    assert abs(sa["final_cost"] - sb["final_cost"]) <= 1e-12 * abs(sa["final_cost"])
    assert sb["h2d_bytes"] == sa["h2d_bytes"]
    ka = make_ka_problem(seed=4)[0]
    kb = ka.copy()
    off()
    s1 = _engine.ka_run(ka, ic, _capi.default_ka_options())
    on(30000)
    s2 = _engine.ka_run(kb, ic, _capi.default_ka_options())
    assert s1["final_cost"] == pytest.approx(s2["final_cost"], rel=1e-12)
    assert np.abs(ka.keypoints - kb.keypoints).max() < 1e-12
