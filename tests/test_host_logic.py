"""CPU-side checks of the product library (no compute kernels are launched): the C-ABI loads and
exports every symbol include/pxr.h declares, the host-side integer algorithms are bit-exact
against fixtures generated from the reference's own graph.cc, and compute entry points refuse to
run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import HAS_GPU, ROOT
from pixsfm._pixsfm import _capi

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_every_declared_symbol_is_exported():
    lib = _capi.load_lib()
    hdr = open(os.path.join(ROOT, "include", "pxr.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(pxr_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 25
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


@pytest.mark.skipif(HAS_GPU, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_device():
    lib = _capi.load_lib()
    h = C.c_void_p()
    rc = lib.pxr_ctx_create(-1, C.byref(h))
    assert rc == _capi.PXR_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.pxr_last_error()
    with pytest.raises(_capi.PxrError):
        _capi.Context(-1)


def test_graph_labels_bit_exact_vs_reference_fixture():
    lib = _capi.load_lib()
    z = np.load(os.path.join(GOLD, "graph_ref.npz"))
    for g in range(int(z["n_graphs"])):
        ni = z["g%d_node_image" % g]; es = z["g%d_es" % g]; ed = z["g%d_ed" % g]; sim = z["g%d_sim" % g]
        n = len(ni)
        tl = np.zeros(n, np.int64); sc = np.zeros(n); rt = np.zeros(n, np.uint8)
        assert lib.pxr_graph_track_labels(C.c_int64(n), _p(ni), C.c_int64(len(es)), _p(es), _p(ed), _p(sim), _p(tl)) == 0
        assert lib.pxr_graph_score_labels(C.c_int64(n), C.c_int64(len(es)), _p(es), _p(ed), _p(sim), _p(tl), _p(sc)) == 0
        assert lib.pxr_graph_root_labels(C.c_int64(n), _p(tl), _p(sc), _p(rt)) == 0
        assert np.array_equal(tl, z["g%d_track_labels" % g])
        assert np.array_equal(sc, z["g%d_scores" % g])
        assert np.array_equal(rt, z["g%d_is_root" % g])


def test_ka_problem_labels_match_python_reference():
    from collections import Counter
    import sys
    lib = _capi.load_lib()

    def ref_labels(track_labels, max_per_problem):  # keypoint_adjustment/main.py:13-57
        track_count = Counter(track_labels)
        bins = []; t2p = [-1] * len(track_count); start = 0; last_v = sys.maxsize
        for k, v in track_count.most_common():
            if v < last_v:
                start = 0; last_v = v
            found = False
            if v < max_per_problem:
                for i in range(start, len(bins)):
                    if bins[i] + v <= max_per_problem:
                        bins[i] += v; t2p[k] = i; found = True; start = i
                        break
            if not found:
                t2p[k] = len(bins); start = len(bins); bins.append(v)
        return [t2p[v] for v in track_labels], bins

    rng = np.random.default_rng(2)
    for trial in range(6):
        n_tracks = int(rng.integers(3, 300))
        labels = np.repeat(np.arange(n_tracks), rng.integers(1, 14, n_tracks))
        rng.shuffle(labels)
        mp = [50, 10, 13, 50, 7, 50][trial]
        exp, bins = ref_labels(labels.tolist(), mp)
        out = np.zeros(len(labels), np.int32); nb = C.c_int32()
        assert lib.pxr_ka_problem_labels(C.c_int64(len(labels)), _p(labels.astype(np.int64)), mp, _p(out), C.byref(nb)) == 0
        assert nb.value == len(bins) and out.tolist() == exp


def test_shard_points_plan():
    lib = _capi.load_lib()
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 20, 1000)
    obs_pt = np.repeat(np.arange(1000, dtype=np.int64), lens)
    for world in (1, 2, 3, 8):
        pb = np.zeros(world + 1, np.int64); ob = np.zeros(world + 1, np.int64)
        assert lib.pxr_shard_points(C.c_int64(1000), C.c_int64(len(obs_pt)), _p(obs_pt), world, _p(pb), _p(ob)) == 0
        assert pb[0] == 0 and pb[-1] == 1000 and ob[0] == 0 and ob[-1] == len(obs_pt)
        assert np.all(np.diff(pb) >= 0) and np.all(np.diff(ob) >= 0)
        for r in range(world):  # a rank owns whole points: its obs range is exactly its points' observations
            sel = (obs_pt >= pb[r]) & (obs_pt < pb[r + 1])
            assert sel.sum() == ob[r + 1] - ob[r]
        assert np.diff(ob).max() - np.diff(ob).min() <= 2 * lens.max() + 1
    bad = obs_pt[::-1].copy()
    pb = np.zeros(3, np.int64); ob = np.zeros(3, np.int64)
    assert lib.pxr_shard_points(C.c_int64(1000), C.c_int64(len(bad)), _p(bad), 2, _p(pb), _p(ob)) == _capi.PXR_ERR_INVALID_ARGUMENT


def test_default_option_blocks_match_reference_python_defaults():
    lib = _capi.load_lib()
    o = _capi.SolverOptions()
    lib.pxr_default_ba_options(C.byref(o))
    ref = _capi.default_ba_options()
    for f, _t in _capi.SolverOptions._fields_:
        assert getattr(o, f) == getattr(ref, f), f
    assert o.loss_type == 1 and o.loss_scale == 0.25 and o.max_num_iterations == 100 and o.use_inner_iterations == 1
    lib.pxr_default_ka_options(C.byref(o))
    assert o.parameter_tolerance == 1e-5 and o.use_inner_iterations == 0


def test_interrupt_callback_and_sigint_polling(monkeypatch):
    """pxr_set_interrupt_callback / pxr_poll_interrupt (host only): the mirror registers a callback that reports a
    pending Ctrl-C the way the reference's PyInterrupt does; a solve that sees it stops with PXR_ERR_INTERRUPTED,
    which the mirror raises as KeyboardInterrupt."""
    import ctypes as C
    import signal
    import threading
    from pixsfm._pixsfm import _capi
    lib = _capi.load_lib()
    if threading.current_thread() is not threading.main_thread():
        pytest.skip("signal handlers run in the main thread only")
    assert lib.pxr_poll_interrupt() == 0                      # nothing pending

    def sigint_arrives_now():
        # Ctrl-C while the library runs: the signal becomes deliverable inside the callback, Python's handler raises there
        signal.pthread_sigmask(signal.SIG_UNBLOCK, {signal.SIGINT})
        for _ in range(1000):
            pass
        return 0
    signal.pthread_sigmask(signal.SIG_BLOCK, {signal.SIGINT})
    try:
        signal.pthread_kill(threading.main_thread().ident, signal.SIGINT)     # stays pending: this thread blocks it
        monkeypatch.setattr(_capi, "_check_signals", sigint_arrives_now)
        assert lib.pxr_poll_interrupt() == 1                  # reported to the library, not leaked as an exception here
    finally:
        signal.pthread_sigmask(signal.SIG_UNBLOCK, {signal.SIGINT})
        monkeypatch.undo()
    assert lib.pxr_poll_interrupt() == 0                      # consumed
    monkeypatch.setattr(_capi, "_check_signals", lambda: -1)  # PyErr_CheckSignals reporting a raised handler
    assert lib.pxr_poll_interrupt() == 1
    monkeypatch.undo()
    hits = []
    cb = C.CFUNCTYPE(C.c_int, C.c_void_p)(lambda user: hits.append(user) or 1)
    try:
        assert lib.pxr_set_interrupt_callback(cb, C.c_void_p(7)) == 0
        assert lib.pxr_poll_interrupt() == 1 and hits == [7]
        assert lib.pxr_set_interrupt_callback(None, None) == 0 and lib.pxr_poll_interrupt() == 0
    finally:
        lib.pxr_set_interrupt_callback(_capi._SIGNAL_POLL, None)
    with pytest.raises(KeyboardInterrupt):
        _capi.check(_capi.PXR_ERR_INTERRUPTED)


def test_device_memory_estimate():
    """pxr_ba_estimate_device_bytes: host-side count of what a BA solve takes on the device"""
    from pixsfm._pixsfm import _capi, _engine
    from pixsfm.util import synthetic
    prob, _ = synthetic.make_ba_scene(n_cams=6, n_points=50, track_len=4, channels=16, seed=1)
    prob.refs = np.zeros((50, 16))
    e = _engine.ba_estimate_device_bytes(prob, _capi.default_ba_options())
    n_obs, K = prob.n_obs, 4                                         # SIMPLE_RADIAL: 4 intrinsics
    assert e["patches"] == n_obs * 16 * 16 * 16 * 2                  # the fp16 slab, exactly
    pairs = 50 * (4 * 5 // 2)
    per_obs = 2 * (2 + 8 + 2 * (9 + K)) * 8 + 2 * (6 + K) * 3 * 8 + (6 + K) * 4 + 4 + 20
    state = n_obs * per_obs + pairs * 8 + 50 * ((9 + 3 + 6) * 8 + 48 + 17) + n_obs * 24 + 50 * 16 * 8
    assert e["state"] == state
    nc = 6 * 6 + K * len(prob.cam_model)
    assert e["reduced_system"] == 2 * nc * nc * 8 and e["total"] == e["patches"] + e["state"] + e["reduced_system"]
    # resident patches cost nothing more; BASELINE configs[2] sizes: the slab dominates (32.8 GB), the rest is < 1 GB
    big = dict(n_obs=500000, n_pts=50000, L=10)
    obs_pt = np.repeat(np.arange(big["n_pts"], dtype=np.int64), big["L"])
    import ctypes as C
    d = prob.desc()
    d.n_obs, d.n_points, d.n_patches, d.n_images = big["n_obs"], big["n_pts"], big["n_obs"], 200
    d.obs_pt = obs_pt.ctypes.data_as(C.c_void_p); d.channels = 128; d.obs_patch = None; d.refs = None
    out = [C.c_double() for _ in range(3)]
    lib = _capi.load_lib()
    assert lib.pxr_ba_estimate_device_bytes(C.byref(d), None, *[C.byref(o) for o in out]) == 0
    assert out[0].value == 500000 * 16 * 16 * 128 * 2 == 32768000000.0
    assert 0.4e9 < out[1].value < 1.0e9
    d.patches_on_device = 1
    assert lib.pxr_ba_estimate_device_bytes(C.byref(d), None, C.byref(out[0]), None, None) == 0 and out[0].value == 0
    # ITERATIVE_SCHUR at config-5 camera counts: block-sparse, far below the dense 2 x 12.8 GB
    d.n_images = 5000
    so = _capi.default_ba_options(linear_solver=3)
    assert lib.pxr_ba_estimate_device_bytes(C.byref(d), C.byref(so), None, None, C.byref(out[2])) == 0
    assert out[2].value < 1e9
    assert lib.pxr_ba_estimate_device_bytes(None, None, None, None, None) != 0
