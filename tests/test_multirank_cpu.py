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
