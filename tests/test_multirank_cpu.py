"""N>1 host logic on CPU (gloo, world_size 2): the point-sharding plan of pxr_shard_points and the additivity the
NCCL path relies on — the all-reduced sum of every rank's camera blocks / gradient / cost equals the
single-process linearisation (checked with the oracle, since no kernel can run here)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "pixel-perfect-sfm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from pixsfm._pixsfm import _capi
    from pixsfm.util import synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob, _ = synthetic.make_ba_scene(n_cams=5, n_points=41, track_len=3, channels=16, seed=33, dtype=np.float64)
    ic = _capi.default_interp(); so = _capi.default_ba_options(use_inner_iterations=0)
    refs, _ = O.refs_compute(prob, ic); prob.refs = refs
    lib = _capi.load_lib()
    pb = np.zeros(world + 1, np.int64); ob = np.zeros(world + 1, np.int64)
    rc = lib.pxr_shard_points(C.c_int64(len(prob.xyz)), C.c_int64(prob.n_obs), prob.obs_pt.ctypes.data_as(C.c_void_p), world,
                              pb.ctypes.data_as(C.c_void_p), ob.ctypes.data_as(C.c_void_p))
    assert rc == 0
    p0, p1, o0, o1 = pb[rank], pb[rank + 1], ob[rank], ob[rank + 1]
    shard = _capi.BAProblem(cam_model=prob.cam_model, cam_params=prob.cam_params, cam_const_mask=prob.cam_const_mask,
                            qvec=prob.qvec, tvec=prob.tvec, img_cam=prob.img_cam, pose_const=prob.pose_const,
                            tvec_const_mask=prob.tvec_const_mask, xyz=prob.xyz[p0:p1], point_const=prob.point_const[p0:p1],
                            obs_img=prob.obs_img[o0:o1], obs_pt=prob.obs_pt[o0:o1] - p0,
                            patches=np.ascontiguousarray(prob.patches[o0:o1]), corner=prob.corner[o0:o1],
                            scale=prob.scale[o0:o1], refs=prob.refs[p0:p1])
    lin = O.ba_linearize(shard, ic, so)
    buf = torch.from_numpy(np.concatenate([lin["Hcc"].ravel(), lin["gc"], [lin["cost"]]]))
    dist.all_reduce(buf)
    full = O.ba_linearize(prob, ic, so)
    nc = full["nc"]
    ok = (np.allclose(buf[:nc * nc].numpy().reshape(nc, nc), full["Hcc"], rtol=1e-11, atol=1e-12)
          and np.allclose(buf[nc * nc:nc * nc + nc].numpy(), full["gc"], rtol=1e-10, atol=1e-12)
          and abs(buf[-1].item() - full["cost"]) <= 1e-12 * full["cost"])
    q.put((rank, bool(ok), int(o1 - o0)))
    dist.destroy_process_group()


def test_sharded_normal_equations_sum_to_the_single_process_ones():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    counts = [n for _, _, n in sorted(res)]
    assert sum(counts) == 41 * 3 and abs(counts[0] - counts[1]) <= 6


def test_ka_problems_shard_without_a_collective():
    """KA shards by whole problems: the plan is a balanced partition, and solving the shards
    independently (here with the oracle as the solver) reproduces the unsharded solve exactly."""
    import oracle_lib as O
    from pixsfm._pixsfm import _capi, _engine
    from ka_util import make_ka_problem, make_query_ka_problem
    prob, _, _ = make_ka_problem(n_images=6, n_tracks=60, track_len=4, channels=16, seed=2, max_per_problem=12)
    ic = _capi.default_interp()
    so = _capi.default_ka_options()
    w = prob.problem_weights()
    assert len(w) == prob.n_problems and w.min() > 0
    per_patch = 16 * 16 * 16 * 2
    kp0 = np.unique(np.concatenate([prob.edge_src[prob.edge_problem == 0], prob.edge_dst[prob.edge_problem == 0]]))
    assert w[0] == len(kp0) * per_patch
    full = prob.copy()
    O.ka_solve(full, ic, so)
    for world in (1, 2, 3):
        plan = _engine.ka_shard_plan(w, world)
        assert plan.min() >= 0 and plan.max() < world
        loads = np.array([w[plan == r].sum() for r in range(world)])
        assert loads.sum() == w.sum() and loads.max() - loads.min() <= w.max()       # LPT bound
        assert np.array_equal(plan, _engine.ka_shard_plan(w, world))                 # deterministic
        merged = prob.copy()
        seen = np.zeros(len(prob.keypoints), int)
        n_edges = 0
        for r in range(world):
            sub, kp_global = merged.shard(plan, r)
            assert sub.n_problems == int(np.sum(plan == r))
            assert np.array_equal(sub.patches, prob.patches[kp_global])
            seen[kp_global] += 1
            n_edges += len(sub.edge_src)
            O.ka_solve(sub, ic, so)
            merged.merge_shard(sub, kp_global)
        assert n_edges == len(prob.edge_src) and seen.max() == 1     # problems share no keypoint
        assert np.array_equal(merged.keypoints, full.keypoints)
    # heaviest-first: a single dominating problem sits alone
    plan = _engine.ka_shard_plan(np.array([1, 1, 10, 1, 1], np.int64), 2)
    assert list(plan) == [1, 1, 0, 1, 1]
    assert list(_engine.ka_shard_plan(np.zeros(0, np.int64), 4)) == []
    # query mode shards the same way (edge_dst indexes the fixed descriptors, which every rank keeps)
    q, _ = make_query_ka_problem(n_images=4, n_tracks=20, track_len=3, channels=16, seed=3)[:2]
    qfull = q.copy()
    O.ka_solve(qfull, ic, so)
    qplan = _engine.ka_shard_plan(q.problem_weights(), 2)
    qm = q.copy()
    for r in range(2):
        sub, kg = qm.shard(qplan, r)
        O.ka_solve(sub, ic, so)
        qm.merge_shard(sub, kg)
    assert np.array_equal(qm.keypoints, qfull.keypoints)


def _ka_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "pixel-perfect-sfm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from pixsfm._pixsfm import _capi, _engine
    from ka_util import make_ka_problem
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = make_ka_problem(n_images=5, n_tracks=40, track_len=4, channels=16, seed=6, max_per_problem=10)[0]
    ic = _capi.default_interp(); so = _capi.default_ka_options()
    plan = _engine.ka_shard_plan(prob.problem_weights(), world)      # every rank computes the same plan, no exchange
    sub, kp_global = prob.shard(plan, rank)
    O.ka_solve(sub, ic, so)                                          # the rank's own problems only
    mine = torch.zeros(len(prob.keypoints), 2, dtype=torch.float64)
    mine[torch.from_numpy(kp_global)] = torch.from_numpy(sub.keypoints)
    owned = torch.zeros(len(prob.keypoints), dtype=torch.int32); owned[torch.from_numpy(kp_global)] = 1
    dist.all_reduce(mine); dist.all_reduce(owned)                    # only to bring the results together for the check
    full = prob.copy()
    O.ka_solve(full, ic, so)
    touched = owned.numpy() > 0
    ok = bool(owned.max().item() == 1 and np.array_equal(mine.numpy()[touched], full.keypoints[touched])
              and np.array_equal(prob.keypoints[~touched], full.keypoints[~touched]))
    q.put((rank, ok, int(sub.n_problems)))
    dist.destroy_process_group()


def test_ka_shards_solve_independently_on_two_ranks():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_ka_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    counts = [n for _, _, n in sorted(res)]
    assert min(counts) >= 1 and abs(counts[0] - counts[1]) <= 2
