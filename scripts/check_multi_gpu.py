"""2+ GPU consistency check of the point-sharded BA (run under torchrun on a multi-GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/check_multi_gpu.py
Every rank solves its shard of ONE seeded problem through pxr_ba_run with the NCCL communicator attached; rank 0
compares the result with the single-process oracle solution of the whole problem (test infrastructure, like tests/).
The last check shards a keypoint adjustment by whole problems (pxr_shard_ka_problems) and compares the same way."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixel-perfect-sfm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from pixsfm._pixsfm import _capi, _engine
    from pixsfm.util import synthetic
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _capi.Context(local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(_capi.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    ctx.init_comm(rank, world, bytes(uid.cpu().numpy().tobytes()))

    ok_all = True
    if rank == 0:
        print("scalar exchange over peer mailboxes (NVLink): %s" % ctx.mailbox_ready(), flush=True)
    for inner, solver, sparse in ((0, 0, 0), (1, 0, 0), (0, 3, 0), (0, 3, 1)):
        if sparse:
            os.environ["PXR_PCG_SPARSE"] = "1"     # implicit block-sparse reduced system, q all-reduced per CG iteration
        else:
            os.environ.pop("PXR_PCG_SPARSE", None)
        prob, _ = synthetic.make_ba_scene(n_cams=40, n_points=480, track_len=5, channels=16, seed=77)
        ic = _capi.default_interp()
        so = _capi.default_ba_options(use_inner_iterations=inner, max_num_iterations=10, linear_solver=solver)
        refs, _ = O.refs_compute(prob, ic); prob.refs = refs
        lib = _capi.load_lib()
        pb = np.zeros(world + 1, np.int64); ob = np.zeros(world + 1, np.int64)
        assert lib.pxr_shard_points(C.c_int64(len(prob.xyz)), C.c_int64(prob.n_obs), prob.obs_pt.ctypes.data_as(C.c_void_p), world,
                                    pb.ctypes.data_as(C.c_void_p), ob.ctypes.data_as(C.c_void_p)) == 0
        p0, p1, o0, o1 = pb[rank], pb[rank + 1], ob[rank], ob[rank + 1]
        shard = _capi.BAProblem(cam_model=prob.cam_model, cam_params=prob.cam_params, cam_const_mask=prob.cam_const_mask,
                                qvec=prob.qvec, tvec=prob.tvec, img_cam=prob.img_cam, pose_const=prob.pose_const,
                                tvec_const_mask=prob.tvec_const_mask, xyz=prob.xyz[p0:p1], point_const=prob.point_const[p0:p1],
                                obs_img=prob.obs_img[o0:o1], obs_pt=prob.obs_pt[o0:o1] - p0,
                                patches=np.ascontiguousarray(prob.patches[o0:o1]), corner=prob.corner[o0:o1],
                                scale=prob.scale[o0:o1], refs=prob.refs[p0:p1])
        h = _engine.BAHandle(shard, ic, so, ctx=ctx)      # set-up collectives (key union) happen here
        ncoll0 = ctx.nccl_collectives()
        s = h.solve()
        ncoll = ctx.nccl_collectives() - ncoll0
        h.read_params()                                   # writes into the shard's arrays
        h.close()
        n_lm = s["num_iterations"] - 1                    # LM iterations after iteration zero
        # every rank must hold the same cameras; points are sharded
        q = torch.from_numpy(np.concatenate([shard.qvec.ravel(), shard.tvec.ravel(), shard.cam_params.ravel()])).cuda()
        qmax = q.clone(); qmin = q.clone()
        dist.all_reduce(qmax, op=dist.ReduceOp.MAX); dist.all_reduce(qmin, op=dist.ReduceOp.MIN)
        same = bool((qmax == qmin).all().item())
        if rank == 0:
            full = prob.copy()
            sr = O.ba_solve(full, ic, so)
            tol = 1e-4 if solver == 3 else 1e-6
            dq = np.abs(shard.qvec - full.qvec).max(); dt = np.abs(shard.tvec - full.tvec).max()
            dx = np.abs(shard.xyz - full.xyz[p0:p1]).max()
            dc = abs(s["final_cost"] - sr["final_cost"]) / sr["final_cost"]
            # ONE NCCL all-reduce per LM iteration (+ one at the end of the solve for the last gradient norm)
            one_collective = (not ctx.mailbox_ready()) or ncoll <= n_lm + 1
            ok = same and dq < tol and dt < tol and dx < tol and dc < 1e-5 and s["num_iterations"] == sr["num_iterations"] and one_collective
            print("inner=%d solver=%d sparse=%d world=%d: ranks identical=%s  |dq|=%.2e |dt|=%.2e |dX|=%.2e  dcost=%.2e  iters %d/%d  "
                  "NCCL collectives %d for %d LM iterations -> %s"
                  % (inner, solver, sparse, world, same, dq, dt, dx, dc, s["num_iterations"], sr["num_iterations"], ncoll, n_lm,
                     "OK" if ok else "MISMATCH"), flush=True)
            ok_all = ok_all and ok
    # ---- keypoint adjustment: whole problems per rank, no collective on the data path; the gather below is only
    # how this check brings the shards' results together
    os.environ.pop("PXR_PCG_SPARSE", None)
    from ka_util import make_ka_problem
    kprob = make_ka_problem(n_images=6, n_tracks=80, track_len=4, channels=128, seed=5, max_per_problem=12)[0]
    kso = _capi.default_ka_options()
    plan = _engine.ka_shard_plan(kprob.problem_weights(), world)
    sub, kp_global = kprob.shard(plan, rank)
    _engine.ka_run(sub, ic, kso, ctx=ctx)
    mine = torch.zeros(len(kprob.keypoints), 2, dtype=torch.float64, device="cuda")
    mine[torch.from_numpy(kp_global).cuda()] = torch.from_numpy(sub.keypoints).cuda()
    dist.all_reduce(mine)       # every keypoint belongs to exactly one shard, the others contribute zeros
    if rank == 0:
        kfull = kprob.copy()
        O.ka_solve(kfull, ic, kso)
        touched = np.zeros(len(kprob.keypoints), bool)
        for r in range(world):
            touched[kprob.shard(plan, r)[1]] = True
        dk = np.abs(mine.cpu().numpy() - kfull.keypoints)[touched].max()
        ok = dk < 1e-5
        print("KA sharded over %d ranks (%d problems): |dkp| vs single-process oracle = %.2e -> %s"
              % (world, kprob.n_problems, dk, "OK" if ok else "MISMATCH"), flush=True)
        ok_all = ok_all and ok
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok_all else "FAIL", flush=True)
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
