"""pxr_solver_options.deterministic = 1: the normal equations are assembled by fixed-order reductions (block mode + chunk
partials summed in chunk order, point blocks observation by observation, scalar sums block by block: csrc/pxr_block.cuh)
instead of fp64 atomics.  Two runs must be bit-identical — iteration records and parameters — and agree with the default
(atomic) path and the oracle to the usual tolerances."""
import numpy as np
import pytest

import oracle_lib as O
from pixsfm._pixsfm import _capi, _engine
from pixsfm.util import synthetic

pytestmark = pytest.mark.gpu


def _scene(**kw):
    args = dict(n_cams=12, n_points=400, track_len=6, channels=32, seed=31)
    args.update(kw)
    prob, _ = synthetic.make_ba_scene(**args)
    ic = _capi.default_interp()
    prob.refs, _ = O.refs_compute(prob, ic)
    return prob, ic


def _bits(summary, p):
    its = [(i["cost"], i["cost_change"], i["step_norm"], i["relative_decrease"], i["trust_region_radius"], i["gradient_max_norm"],
            i["step_is_successful"]) for i in summary["iterations"]]
    return its, p.qvec.tobytes(), p.tvec.tobytes(), p.xyz.tobytes(), p.cam_params.tobytes()


@pytest.mark.parametrize("inner,solver,shared", [(0, 0, False), (1, 0, False), (0, 3, False), (1, 0, True)])
def test_two_deterministic_runs_are_bit_identical(inner, solver, shared):
    prob, ic = _scene(shared_camera=shared)
    so = _capi.default_ba_options(max_num_iterations=10, use_inner_iterations=inner, linear_solver=solver, deterministic=1)
    runs = []
    for _ in range(3):
        p = prob.copy()
        s = _engine.ba_run(p, ic, so)
        runs.append(_bits(s, p))
    assert runs[0] == runs[1] == runs[2]
    # and it is the same solve as the default path / the oracle
    so0 = _capi.default_ba_options(max_num_iterations=10, use_inner_iterations=inner, linear_solver=solver)
    p0, pd, pc = prob.copy(), prob.copy(), prob.copy()
    s0 = _engine.ba_run(p0, ic, so0)
    sd = _engine.ba_run(pd, ic, so)
    sc = O.ba_solve(pc, ic, so0)
    assert len(sd["iterations"]) == len(s0["iterations"]) == len(sc["iterations"])
    assert abs(sd["final_cost"] - s0["final_cost"]) <= 1e-9 * s0["final_cost"]
    assert abs(sd["final_cost"] - sc["final_cost"]) <= 1e-6 * sc["final_cost"]
    assert np.abs(pd.xyz - p0.xyz).max() < 1e-8 and np.abs(pd.xyz - pc.xyz).max() < 1e-6


def test_long_tracks_and_window_residency_do_not_break_reproducibility():
    """tracks longer than a warp (several partial sums per point) and the window upload (refetches) in the same run"""
    prob, ic = _scene(n_cams=40, n_points=60, track_len=36, channels=16, rot_sigma_deg=0.06, pt_sigma=0.012)
    so = _capi.default_ba_options(max_num_iterations=8, use_inner_iterations=1, deterministic=1)
    a, b = prob.copy(), prob.copy()
    sa, sb = _engine.ba_run(a, ic, so), _engine.ba_run(b, ic, so)
    assert _bits(sa, a) == _bits(sb, b) and sa["resident_window"] == 8
