"""`pixsfm._pixsfm._residuals` (reference: residuals/bindings.cc:14-30): the cost-functor factories.
CPU: construction rules, parameter-block layout, the tangent -> ambient quaternion map.
GPU: evaluate() against the oracle's forward-mode Jets over the reference's functor (oracle/orc_ba.h::EvalBlockJets,
restating residuals/src/feature_reference.h:98-137 as ceres::AutoDiffCostFunction evaluates it)."""
import numpy as np
import pytest

import oracle_lib as O
from pixsfm._pixsfm import _capi, _residuals
from pixsfm._pixsfm._features import FeaturePatch
from pixsfm.util import synthetic


def _patch(channels=128, dtype=np.float16, seed=0):
    rng = np.random.default_rng(seed)
    data = rng.normal(0, 1, (16, 16, channels)).astype(dtype)
    return FeaturePatch(data, corner=(492, 493), scale=(1.0, 1.0))


def test_factories_follow_the_reference_signatures_and_checks():
    p = _patch()
    ic = {"l2_normalize": True}
    f = _residuals.FeatureReferenceCostFunctor(2, p, np.zeros((1, 128)), ic)
    assert f.parameter_block_sizes() == [4, 3, 3, 4] and f.num_residuals() == 128      # SIMPLE_RADIAL: 4 parameters
    g = _residuals.FeatureReferenceConstantPoseCostFunctor(1, [1, 0, 0, 0], [0, 0, 1], p, np.zeros((1, 128)), ic)
    assert g.parameter_block_sizes() == [3, 4] and g.num_residuals() == 128            # PINHOLE
    with pytest.raises(ValueError):                       # THROW_CHECK_EQ(reference_descriptor.cols(), channels)
        _residuals.FeatureReferenceCostFunctor(2, p, np.zeros((1, 64)), ic)
    with pytest.raises(ValueError, match="Unsupported dimensions"):      # n_nodes != 1
        _residuals.FeatureReferenceCostFunctor(2, p, np.zeros((2, 128)), {"nodes": [[0, 0], [1, 1]]})
    with pytest.raises(ValueError, match="Unsupported dimensions"):
        _residuals.FeatureMetricCostFunctor(None, p, None, p, ic)
    h = _residuals.GeometricCostFunctor(2, [10.0, 20.0])
    assert h.parameter_block_sizes() == [4, 3, 3, 4] and h.num_residuals() == 2
    assert _residuals.GeometricConstantPoseCostFunctor(0, [1, 0, 0, 0], [0, 0, 0], [1, 2]).parameter_block_sizes() == [3, 3]
    import pixsfm.residuals as R
    assert R.FeatureReferenceCostFunctor is _residuals.FeatureReferenceCostFunctor


def test_tangent_to_ambient_map_matches_finite_differences_of_quaternion_plus():
    rng = np.random.default_rng(3)
    q = rng.normal(size=4) * 1.7                      # not unit: the functor normalises
    E = _residuals._tangent_to_ambient(q)
    assert np.abs(E @ q).max() < 1e-14                # the radial direction has no effect
    # an ambient step dq equals, to first order, the manifold step delta = E dq on the normalised quaternion
    dq = rng.normal(size=4) * 1e-6
    delta = E @ dq
    qn = q / np.linalg.norm(q)
    plus = synthetic.quat_mul(np.concatenate([[1.0], delta]), qn)
    plus /= np.linalg.norm(plus)
    target = (q + dq) / np.linalg.norm(q + dq)
    assert np.abs(plus - target).max() < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("model,channels,dtype", [(2, 128, np.float16), (1, 16, np.float16), (4, 128, np.float32), (2, 1, np.float64)])
def test_feature_reference_functors_match_the_oracle_jets(model, channels, dtype):
    rng = np.random.default_rng(model * 10 + channels)
    p = _patch(channels, dtype, seed=channels)
    k = _capi.CAMERA_NUM_PARAMS[model]
    cam = np.array([1200.0, 1190.0, 500.0, 500.0, 0.02, -0.01, 0.001, 0.002][:k]) if model == 4 else \
        (np.array([1200.0, 500.0, 500.0, 0.03]) if model == 2 else np.array([1200.0, 1190.0, 500.0, 500.0]))
    q = np.array([0.98, 0.05, -0.03, 0.02]) * 1.3
    t = np.array([0.01, -0.02, 10.0])
    X = np.array([0.02, 0.03, 0.1])
    ref = rng.normal(size=(1, channels)); ref /= np.linalg.norm(ref)
    ic = {"l2_normalize": channels > 1}
    # the oracle over the same one-observation problem
    def oracle(with_ref):
        prob = _capi.BAProblem(cam_model=[model], cam_params=[cam], cam_const_mask=[0], qvec=[q], tvec=[t], img_cam=[0],
                               pose_const=[0], tvec_const_mask=[0], xyz=[X], point_const=[0], obs_img=[0], obs_pt=[0],
                               patches=np.ascontiguousarray(p.data)[None], corner=[p.corner], scale=[p.scale],
                               refs=ref if with_ref else None)
        r, J = O.ba_block_jacobians(prob, _capi.default_interp(channels > 1, False))
        return r[0], J[0]
    # non-constant pose: the reference's factory drops the descriptor (feature_reference.h:269-270)
    f = _residuals.FeatureReferenceCostFunctor(model, p, ref, ic)
    r, (Jq, Jt, JX, Jc) = f.evaluate(q, t, X, cam)
    ro, Jo = oracle(False)
    scale = max(1.0, np.abs(Jo).max())
    assert np.abs(r - ro).max() < 1e-12
    assert np.abs(Jq - Jo[:, 0:4]).max() < 1e-9 * scale and np.abs(Jt - Jo[:, 4:7]).max() < 1e-9 * scale
    assert np.abs(JX - Jo[:, 7:10]).max() < 1e-9 * scale and np.abs(Jc - Jo[:, 10:10 + k]).max() < 1e-9 * scale
    assert np.abs(f(q, t, X, cam) - ro).max() < 1e-12
    # constant pose: residual = f - reference, blocks (xyz, cam)
    g = _residuals.FeatureReferenceConstantPoseCostFunctor(model, q, t, p, ref, ic)
    r2, (JX2, Jc2) = g.evaluate(X, cam)
    ro2, Jo2 = oracle(True)
    assert np.abs(r2 - ro2).max() < 1e-12
    assert np.abs(JX2 - Jo2[:, 7:10]).max() < 1e-9 * scale and np.abs(Jc2 - Jo2[:, 10:10 + k]).max() < 1e-9 * scale


@pytest.mark.gpu
def test_geometric_functor_is_the_reprojection_error_with_its_jacobian():
    q = np.array([0.98, 0.05, -0.03, 0.02]); t = np.array([0.01, -0.02, 10.0]); X = np.array([0.02, 0.03, 0.1])
    cam = np.array([1200.0, 500.0, 500.0, 0.03])
    h = _residuals.GeometricCostFunctor(2, [501.0, 502.0])
    r, (Jq, Jt, JX, Jc) = h.evaluate(q, t, X, cam)
    def proj(q_, t_, X_, c_):
        out = np.zeros(2)
        O.lib().orc_world_to_pixel(2, O.p(np.ascontiguousarray(c_, np.float64)), O.p(np.ascontiguousarray(q_, np.float64)),
                                   O.p(np.ascontiguousarray(t_, np.float64)), O.p(np.ascontiguousarray(X_, np.float64)), O.p(out))
        return out
    assert np.abs(r - (proj(q, t, X, cam) - [501.0, 502.0])).max() < 1e-9
    eps = 1e-6
    for J, idx, n in ((Jq, 0, 4), (Jt, 1, 3), (JX, 2, 3), (Jc, 3, 4)):
        for c in range(n):
            a = [q.copy(), t.copy(), X.copy(), cam.copy()]; b = [q.copy(), t.copy(), X.copy(), cam.copy()]
            a[idx][c] += eps; b[idx][c] -= eps
            fd = (proj(*a) - proj(*b)) / (2 * eps)
            assert np.abs(J[:, c] - fd).max() < 1e-5 * max(1.0, np.abs(fd).max())
    g = _residuals.GeometricConstantPoseCostFunctor(2, q, t, [501.0, 502.0])
    r2, (JX2, Jc2) = g.evaluate(X, cam)
    assert np.abs(r2 - r).max() < 1e-12 and np.abs(JX2 - JX).max() < 1e-12 and np.abs(Jc2 - Jc).max() < 1e-12
