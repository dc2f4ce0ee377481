"""Pins the CPU oracle (oracle/) against the reference's own known-answer tests and golden
vectors.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from pixsfm._pixsfm import _capi

GOLD = os.path.join(os.path.dirname(__file__), "golden")
p = O.p


# --- (i) polynomial exactness, reference pixsfm/base/src/interpolation_test.cc:21-58,93-185 ------
def _poly_case(coeff, C_):
    rows = cols = 10
    grid = np.zeros((rows, cols, C_))
    def F(r, c):
        x = np.array([r, c, 1.0]); return x @ coeff @ x
    def dFdr(r, c):
        x = np.array([r, c, 1.0]); return (coeff[0] + coeff[:, 0]) @ x
    def dFdc(r, c):
        x = np.array([r, c, 1.0]); return (coeff[1] + coeff[:, 1]) @ x
    for r in range(rows):
        for c in range(cols):
            for dim in range(C_):
                grid[r, c, dim] = (dim * dim + 1) * F(r, c)
    for j in range(0, 100, 7):
        r = 1.0 + 7.0 / 99 * j
        for k in range(0, 100, 7):
            c = 1.0 + 7.0 / 99 * k
            f, dr, dc = O.pixel_interp(grid, r, c, l2_normalize=False)
            for dim in range(C_):
                s = dim * dim + 1
                assert abs(f[dim] - s * F(r, c)) < 1e-8
                assert abs(dr[dim] - s * dFdr(r, c)) < 1e-8
                assert abs(dc[dim] - s * dFdc(r, c)) < 1e-8


@pytest.mark.parametrize("name", ["zero", "deg00", "deg01", "deg10", "deg11"])
def test_bicubic_polynomial_exactness(name):
    coeff = np.zeros((3, 3))
    if name != "zero":
        coeff[2, 2] = 1.0
    if name == "deg01":
        coeff[0, 2] = coeff[2, 0] = 0.1
    if name in ("deg10", "deg11"):
        coeff[0, 1] = coeff[1, 0] = 0.1
    if name == "deg11":
        coeff[0, 2] = coeff[2, 0] = 0.2
    for C_ in (1, 2, 3):
        _poly_case(coeff, C_)


# --- (ii) L2 normalisation, interpolation_test.cc:187-207 ------------------------------------------
def test_l2_normalize_fixed_grid():
    values = np.array([1.0, 5.0, 2.0, 10.0, 2.0, 6.0, 3.0, 5.0, 1.0, 2.0, 2.0, 2.0, 2.0, 2.0, 3.0, 1.0])
    grid = values.reshape(2, 4, 2)
    for r, c in ((0.5, 2.5), (1.5, 1.5), (0.0, 3.0)):
        f, dr, dc = O.pixel_interp(grid, r, c, l2_normalize=True)
        assert abs(1.0 - f[0] ** 2 - f[1] ** 2) < 1e-10
        # normalised derivative is tangent to the unit sphere
        assert abs(f @ dr) < 1e-10 and abs(f @ dc) < 1e-10


# --- (iv) SIMD path vs ceres bicubic, interpolation_test.cc:327-364 (tolerance 1e-5) ------------------
@pytest.mark.parametrize("dtype", [np.float16, np.float32, np.float64])
def test_simd_bicubic_similar_to_ceres(dtype):
    rng = np.random.default_rng(3)
    grid = rng.uniform(-1, 1, (10, 10, 128)).astype(dtype)
    worst = 0.0
    for r in range(0, 100, 3):
        for c in range(0, 100, 3):
            f, dr, dc = O.bicubic_ceres(grid, r / 10.0, c / 10.0)
            f2, dr2, dc2 = O.pixel_interp(grid, r / 10.0, c / 10.0, l2_normalize=False)
            worst = max(worst, np.abs(f - f2).max(), np.abs(dr - dr2).max(), np.abs(dc - dc2).max())
    assert worst < 1e-5


# --- (iii) Jet chain rule: J = dfdr * dr/dx + dfdc * dc/dx checked by finite differences -------------
def _small_scene():
    from pixsfm.util import synthetic
    prob, gt = synthetic.make_ba_scene(n_cams=4, n_points=12, track_len=3, channels=16, seed=5,
                                       dtype=np.float64, noise=0.0)
    ic = _capi.default_interp()
    so = _capi.default_ba_options(use_inner_iterations=0)
    refs, _ = O.refs_compute(prob, ic)
    prob.refs = refs * 0.9 + 0.01  # keep residuals away from zero
    return prob, ic, so


def test_jet_jacobians_match_finite_differences():
    prob, ic, so = _small_scene()
    lin = O.ba_linearize(prob, ic, so)
    nc, nl, pose_off, intr_off, point_off = O.ba_layout(prob)
    base = lin["cost"]
    eps = 1e-6

    def cost_of(q):
        return O.ba_evaluate(q, ic, so)["cost"]

    # points
    for pt in (0, 5, 11):
        for a in range(3):
            q = prob.copy(); q.xyz[pt, a] += eps
            q2 = prob.copy(); q2.xyz[pt, a] -= eps
            fd = (cost_of(q) - cost_of(q2)) / (2 * eps)
            assert abs(fd - lin["gp"][pt, a]) < 1e-5 * max(1.0, abs(fd))
    # translations (image 2: fully variable)
    img = 2
    for a in range(3):
        q = prob.copy(); q.tvec[img, a] += eps
        q2 = prob.copy(); q2.tvec[img, a] -= eps
        fd = (cost_of(q) - cost_of(q2)) / (2 * eps)
        assert abs(fd - lin["gc"][pose_off[img] + 3 + a]) < 1e-5 * max(1.0, abs(fd))
    # rotation through the quaternion manifold
    for a in range(3):
        d = np.zeros(3); d[a] = eps
        qp = np.zeros(4); qm = np.zeros(4)
        O.lib().orc_quaternion_plus(p(prob.qvec[img]), p(d), p(qp))
        O.lib().orc_quaternion_plus(p(prob.qvec[img]), p(-d), p(qm))
        q = prob.copy(); q.qvec[img] = qp
        q2 = prob.copy(); q2.qvec[img] = qm
        fd = (cost_of(q) - cost_of(q2)) / (2 * eps)
        assert abs(fd - lin["gc"][pose_off[img] + a]) < 1e-5 * max(1.0, abs(fd))
    # intrinsics: SIMPLE_RADIAL with pp constant -> local params are (f, k)
    cam = 1
    for la, idx in enumerate((0, 3)):
        h = eps * (100.0 if idx == 0 else 1.0)
        q = prob.copy(); q.cam_params[cam, idx] += h
        q2 = prob.copy(); q2.cam_params[cam, idx] -= h
        fd = (cost_of(q) - cost_of(q2)) / (2 * h)
        assert abs(fd - lin["gc"][intr_off[cam] + la]) < 1e-5 * max(1.0, abs(fd))


def test_model_cost_change_and_schur_consistency():
    prob, ic, so = _small_scene()
    lin = O.ba_linearize(prob, ic, so, radius=1e4)
    assert lin["model_cost_change"] > 0
    # S * delta_c == rhs
    nc = lin["nc"]
    S = lin["S"]; S = np.tril(S) + np.tril(S, -1).T
    assert np.allclose(S @ lin["delta"][:nc], lin["rhs"], rtol=1e-8, atol=1e-10)


def test_iterative_schur_pcg_restatement():
    """Ceres ITERATIVE_SCHUR + SCHUR_JACOBI as restated in oracle/orc_ba.h (SchurJacobiPCG): the inexact steps
    must converge to the exact solver's optimum, with CG iteration counts reported and bounded."""
    prob, ic, so = _small_scene()
    pe = prob.copy(); pi = prob.copy()
    so_e = _capi.default_ba_options(use_inner_iterations=0, max_num_iterations=15)
    so_i = _capi.default_ba_options(use_inner_iterations=0, max_num_iterations=15, linear_solver=3)
    se = O.ba_solve(pe, ic, so_e); si = O.ba_solve(pi, ic, so_i)
    assert abs(si["final_cost"] - se["final_cost"]) <= 1e-3 * se["final_cost"]
    its = [i["linear_solver_iterations"] for i in si["iterations"][1:]]
    assert its and min(its) >= 1 and max(its) <= so_i.max_linear_solver_iterations
    assert all(i["linear_solver_iterations"] == 0 for i in se["iterations"][:1])
    # (parameters are not compared: this scene has scale-gauge freedom, inexact steps drift along it)


# --- (v) IRLS, reference pixsfm/base/src/irls_optim.h:23-71 vs a numpy restatement -------------------
@pytest.mark.parametrize("C_,n", [(128, 10), (128, 100), (3, 10), (3, 1000)])
def test_irls_matches_numpy(C_, n):
    rng = np.random.default_rng(C_ + n)
    D = rng.uniform(-1, 1, (n, C_))
    mean = np.zeros(C_)
    O.lib().orc_robust_mean_irls(p(D), n, C_, 1, C.c_double(0.25), 100, 0, p(mean))
    w = np.ones(n)
    for _ in range(100):
        w = w / w.sum()
        m = (D * w[:, None]).sum(0)
        s = ((D - m) ** 2).sum(1)
        rho = 0.0625 * np.log1p(s / 0.0625)
        w = 1.0 / rho
    assert np.allclose(mean, m, atol=1e-8)


# --- (vi) projection, reference pixsfm/base/src/projection_test.cc:24-38 ----------------------------
def test_world_to_pixel_known_values():
    xy = np.zeros(2)
    q = np.array([1.0, 0, 0, 0]); t = np.zeros(3); X = np.array([0.0, 0.0, 1.0])
    for model, params in ((0, [655.123, 386.123, 511.123]), (3, [651.123, 386.123, 511.123, 0.05, 0.03])):
        cp = np.zeros(12); cp[:len(params)] = params
        O.lib().orc_world_to_pixel(model, p(cp), p(q), p(t), p(X), p(xy))
        assert np.allclose(xy, [386.123, 511.123], atol=1e-12)
    # off-axis point, RADIAL: numpy restatement of the colmap formula
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.normal(size=4); t = rng.normal(size=3) + np.array([0, 0, 5.0]); X = rng.uniform(-1, 1, 3)
        cp = np.zeros(12); cp[:5] = [651.123, 386.123, 511.123, 0.05, 0.03]
        O.lib().orc_world_to_pixel(3, p(cp), p(q), p(t), p(X), p(xy))
        from pixsfm.util.synthetic import quat_to_R
        pc = quat_to_R(q) @ X + t
        u, v = pc[0] / pc[2], pc[1] / pc[2]
        r2 = u * u + v * v
        rad = cp[3] * r2 + cp[4] * r2 * r2
        assert np.allclose(xy, [cp[0] * (u + u * rad) + cp[1], cp[0] * (v + v * rad) + cp[2]], atol=1e-9)


# --- (vi-b) the camera models on the parameter vectors and the [-0.5, 0.5]^2 grid of the reference's undistortion_test.cc:
# 67-101 (SIMPLE_PINHOLE, PINHOLE, SIMPLE_RADIAL, RADIAL; WorldToImage then back to the normalised point to 1e-6, :15-27).
# The inverse is a Newton iteration written here; the oracle and the Python restatement must both be the forward map.
@pytest.mark.parametrize("model,params", [
    (0, [655.123, 386.123, 511.123]), (1, [651.123, 655.123, 386.123, 511.123]),
    (2, [651.123, 386.123, 511.123, 0.0]), (2, [651.123, 386.123, 511.123, 0.1]),
    (3, [651.123, 386.123, 511.123, 0.0, 0.0]), (3, [651.123, 386.123, 511.123, 0.1, 0.0]),
    (3, [651.123, 386.123, 511.123, 0.05, 0.0]), (3, [651.123, 386.123, 511.123, 0.05, 0.03])])
def test_world_to_image_round_trip_on_the_reference_grid(model, params):
    from pixsfm.util import cameras
    cp = np.zeros(12); cp[:len(params)] = params
    q = np.array([1.0, 0, 0, 0]); t = np.zeros(3); xy = np.zeros(2)
    fx = params[0]; fy = params[1] if model == 1 else params[0]
    cx, cy = (params[2], params[3]) if model == 1 else (params[1], params[2])
    k1 = params[3] if model in (2, 3) else 0.0
    k2 = params[4] if model == 3 else 0.0
    for u0 in np.arange(-0.5, 0.5001, 0.1):
        for v0 in np.arange(-0.5, 0.5001, 0.1):
            O.lib().orc_world_to_pixel(model, p(cp), p(q), p(t), p(np.array([u0, v0, 1.0])), p(xy))
            py = cameras.world_to_image(model, np.asarray(params), q, t, np.array([[u0, v0, 1.0]]))[0]
            assert np.abs(py - xy).max() < 1e-10
            # back: undo the affine part, then solve d(u, v) = (u, v)(1 + k1 r^2 + k2 r^4) for (u, v)
            xd, yd = (xy[0] - cx) / fx, (xy[1] - cy) / fy
            u, v = xd, yd
            for _ in range(50):
                r2 = u * u + v * v
                s = 1 + k1 * r2 + k2 * r2 * r2
                ds = 2 * k1 + 4 * k2 * r2
                J = np.array([[s + u * u * ds, u * v * ds], [u * v * ds, s + v * v * ds]])
                step = np.linalg.solve(J, np.array([u * s - xd, v * s - yd]))
                u, v = u - step[0], v - step[1]
            assert abs(u - u0) < 1e-6 and abs(v - v0) < 1e-6


def test_loss_functions_derivatives():
    rho = np.zeros(3); rp = np.zeros(3); rm = np.zeros(3)
    for t in range(5):
        for s in (0.0, 0.01, 0.05, 0.3, 2.0):
            O.lib().orc_loss(t, C.c_double(0.25), C.c_double(1.0), C.c_double(s), p(rho))
            h = 1e-6
            O.lib().orc_loss(t, C.c_double(0.25), C.c_double(1.0), C.c_double(s + h), p(rp))
            O.lib().orc_loss(t, C.c_double(0.25), C.c_double(1.0), C.c_double(max(s - h, 0)), p(rm))
            if s > 0 and not (t == 2 and abs(s - 0.0625) < 1e-3):
                assert abs((rp[0] - rm[0]) / (2 * h) - rho[1]) < 1e-4
    O.lib().orc_loss(1, C.c_double(0.25), C.c_double(1.0), C.c_double(0.0625), p(rho))
    assert np.allclose(rho, [0.0625 * np.log(2), 0.5, -4.0])


# --- (viii) cost maps, reference costmap_extractor.h:230-358 vs an independent numpy restatement ------
@pytest.mark.parametrize("dtype", [np.float16, np.float64])
def test_costmap_extraction_matches_numpy(dtype):
    from pixsfm.util import synthetic
    prob, gt = synthetic.make_ba_scene(n_cams=4, n_points=12, track_len=3, channels=16, seed=2, dtype=dtype)
    ic = _capi.default_interp()
    prob.refs = O.refs_compute(prob, ic)[0]
    got = O.costmaps_compute(prob, loss_type=1, loss_scale=0.25, as_gradientfield=True, apply_sqrt=False)
    P = prob.patches
    H, W = P.shape[1:3]
    up = np.minimum(np.arange(H) + 1, H - 1); dn = np.maximum(np.arange(H) - 1, 0)
    rt = np.minimum(np.arange(W) + 1, W - 1); lf = np.maximum(np.arange(W) - 1, 0)
    for o in (0, 5, prob.n_obs - 1):
        f = P[o].astype(np.float64)
        # the difference is formed in the patch dtype (Eigen evaluates Map<dtype> - Map<dtype> first)
        dfdr = 0.5 * (P[o][up] - P[o][dn]).astype(dtype).astype(np.float64)
        dfdc = 0.5 * (P[o][:, rt] - P[o][:, lf]).astype(dtype).astype(np.float64)
        r = f - prob.refs[prob.obs_pt[o]]
        s = (r * r).sum(-1)
        b = 0.25 ** 2
        rho0 = b * np.log1p(s / b); rho1 = 1.0 / (1.0 + s / b)
        cost = 0.5 * rho0
        live = cost > 1e-8
        want = np.stack([cost, np.where(live, rho1 * (r * dfdr).sum(-1), 0.0), np.where(live, rho1 * (r * dfdc).sum(-1), 0.0)], -1)
        tol = 2e-3 if dtype == np.float16 else 1e-12
        assert np.allclose(got[o].astype(np.float64), want, rtol=tol, atol=tol * 1e-2 + 1e-15)
    # a patch identical to its reference everywhere has zero cost and (cost <= 1e-8 branch) zero gradient
    prob2 = prob.with_patches(np.ascontiguousarray(np.broadcast_to(prob.refs[prob.obs_pt][:, None, None, :], P.shape).astype(dtype)),
                              refs=prob.refs[:].astype(dtype).astype(np.float64))
    z = O.costmaps_compute(prob2)
    assert np.abs(z.astype(np.float64)).max() == 0.0
