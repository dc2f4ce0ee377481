"""Thin object wrappers over the C-ABI handles (plumbing only; no arithmetic)."""
import ctypes as C

import numpy as np

from . import _capi


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class BAHandle:
    """pxr_ba: a BA problem resident on the device (== FeatureReferenceBundleOptimizer.set_up)."""

    def __init__(self, problem, interp=None, options=None, ctx=None):
        self.ctx = ctx or _capi.default_context()
        self.lib = self.ctx.lib
        self.problem = problem
        self.interp = interp or _capi.default_interp()
        self.options = options or _capi.default_ba_options()
        self._desc = problem.desc()
        self.handle = C.c_void_p()
        _capi.check(self.lib.pxr_ba_create(self.ctx.handle, C.byref(self._desc), C.byref(self.interp),
                                           C.byref(self.options), C.byref(self.handle)))

    def close(self):
        if self.handle:
            self.lib.pxr_ba_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def evaluate(self, residuals=False):
        n = self.problem.n_obs
        out = dict(sq_norm=np.zeros(n), gtr=np.zeros((n, 2)), gtg=np.zeros((n, 3)), xy=np.zeros((n, 2)))
        res = np.zeros((n, self.problem.channels)) if residuals else None
        cost = C.c_double()
        _capi.check(self.lib.pxr_ba_evaluate(self.handle, _p(out["sq_norm"]), _p(out["gtr"]), _p(out["gtg"]),
                                             _p(out["xy"]), _p(res), C.byref(cost)))
        out["cost"] = cost.value
        if residuals:
            out["residuals"] = res
        return out

    def evaluate_jacobians(self):
        """residuals [n,C], grad [n,2,C] (d r/d u, d r/d v), juv [n,2,9+K] (d(u,v)/d(rot3|t3|X3|cam K)), xy [n,2]:
        pxr_ba_evaluate_jacobians, the factored form of every block's ceres-style Jacobian (J = grad^T-rows x juv)"""
        n, Cc = self.problem.n_obs, self.problem.channels
        kmax = int(np.asarray(self.problem.cam_params).shape[1])
        res, grad, xy = np.zeros((n, Cc)), np.zeros((n, 2, Cc)), np.zeros((n, 2))
        juv = np.zeros((n, 2, 9 + kmax)); w = C.c_int32()
        _capi.check(self.lib.pxr_ba_evaluate_jacobians(self.handle, _p(res), _p(grad), _p(juv), C.byref(w), _p(xy)))
        juv = juv.reshape(-1)[: n * 2 * w.value].reshape(n, 2, w.value)
        return dict(residuals=res, grad=grad, juv=juv, xy=xy)

    def debug_linearize(self, nc, nl, radius=1e4, dense=True):
        """dense=False leaves Hcc / S out (the block-sparse and block-mode paths never form them)"""
        npts = len(self.problem.xyz)
        out = dict(gc=np.zeros(nc), Hpp=np.zeros((npts, 3, 3)), gp=np.zeros((npts, 3)), rhs=np.zeros(nc), delta=np.zeros(nl))
        if dense:
            out.update(Hcc=np.zeros((nc, nc)), S=np.zeros((nc, nc)))
        cost = C.c_double(); mcc = C.c_double()
        _capi.check(self.lib.pxr_ba_debug_linearize(self.handle, C.c_double(radius), C.byref(cost), _p(out.get("Hcc")),
                                                    _p(out["gc"]), _p(out["Hpp"]), _p(out["gp"]), _p(out.get("S")),
                                                    _p(out["rhs"]) if dense else None, _p(out["delta"]), C.byref(mcc)))
        out["cost"] = cost.value; out["model_cost_change"] = mcc.value
        return out

    def debug_inner_iterations(self):
        _capi.check(self.lib.pxr_ba_debug_inner_iterations(self.handle))

    def solve(self, capacity=512):
        s = _capi.make_summary(capacity)
        _capi.check(self.lib.pxr_ba_solve(self.handle, C.byref(s)))
        return _capi.summary_to_dict(s)

    def iterate(self, n, capacity=512):
        s = _capi.make_summary(capacity)
        _capi.check(self.lib.pxr_ba_iterate(self.handle, int(n), C.byref(s)))
        return _capi.summary_to_dict(s)

    def kernel_timing(self, enable=-1, which=1, read=True):
        ms = C.c_double(); cnt = C.c_int()
        _capi.check(self.lib.pxr_ba_kernel_timing(self.handle, int(enable), int(which),
                                                  C.byref(ms) if read else None, C.byref(cnt) if read else None))
        return ms.value, cnt.value

    def read_params(self):
        """Copies the current device parameters back into the BAProblem arrays."""
        p = self.problem
        _capi.check(self.lib.pxr_ba_read_params(self.handle, _p(p.cam_params), _p(p.qvec), _p(p.tvec), _p(p.xyz)))

    def reset(self, cam_params, qvec, tvec, xyz):
        """pxr_ba_reset: back to the given parameters, LM trajectory forgotten (patches / work space stay resident)"""
        arrs = [np.ascontiguousarray(a, np.float64) for a in (cam_params, qvec, tvec, xyz)]
        _capi.check(self.lib.pxr_ba_reset(self.handle, *[_p(a) for a in arrs]))

    def time_stage(self, stage, iters):
        ms = C.c_double()
        _capi.check(self.lib.pxr_ba_time_stage(self.handle, int(stage), int(iters), C.byref(ms)))
        return ms.value


def ba_run(problem, interp=None, options=None, ctx=None, capacity=512):
    """pxr_ba_run: upload, solve, write the refined parameters back into `problem` (in place)."""
    ctx = ctx or _capi.default_context()
    interp = interp or _capi.default_interp()
    options = options or _capi.default_ba_options()
    d = problem.desc()
    s = _capi.make_summary(capacity)
    _capi.check(ctx.lib.pxr_ba_run(ctx.handle, C.byref(d), C.byref(interp), C.byref(options), C.byref(s)))
    return _capi.summary_to_dict(s)


class PinnedArray:
    """A numpy array in pinned (page-locked, device-mapped) host memory from pxr_host_alloc_pinned: the patch source
    pxr_ba_run reads tap windows from directly (window residency, csrc/pxr_resident.cuh) and full-rate DMA otherwise.
    `array` is the numpy view; the memory is released by close() / when the object dies."""

    def __init__(self, shape, dtype, ctx=None):
        self.ctx = ctx or _capi.default_context()
        self.array = None
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        self._ptr = C.c_void_p()
        _capi.check(self.ctx.lib.pxr_host_alloc_pinned(C.byref(self._ptr), C.c_size_t(max(nbytes, 1))))
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if self._ptr:
            self.array = None
            self.ctx.lib.pxr_host_free_pinned(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def refs_compute(problem, interp=None, loss_type=1, loss_scale=0.25, iters=100, ctx=None):
    """pxr_refs_compute == ReferenceExtractor.run: -> (refs [n_points,C] f64, src_obs [n_points])"""
    ctx = ctx or _capi.default_context()
    interp = interp or _capi.default_interp()
    d = problem.desc()
    refs = np.zeros((len(problem.xyz), problem.channels))
    src = np.zeros(len(problem.xyz), np.int64)
    s = _capi.make_summary(0)
    _capi.check(ctx.lib.pxr_refs_compute(ctx.handle, C.byref(d), C.byref(interp), int(loss_type), C.c_double(loss_scale),
                                         int(iters), _p(refs), _p(src), C.byref(s)))
    return refs, src


def obs_descriptors(problem, interp=None, ctx=None):
    """pxr_obs_descriptors: [n_obs, C] descriptors at the current projections (Reference.observations)"""
    ctx = ctx or _capi.default_context()
    interp = interp or _capi.default_interp()
    d = problem.desc()
    out = np.zeros((problem.n_obs, problem.channels))
    _capi.check(ctx.lib.pxr_obs_descriptors(ctx.handle, C.byref(d), C.byref(interp), _p(out)))
    return out


def costmaps_compute(problem, interp=None, cfg=None, refs=None, to_host=True, to_device=False, ctx=None):
    """pxr_costmaps_compute == CostMapExtractor.run: reference extraction (unless cfg.compute_refs == 0) and one
    cost patch per observation patch.  -> dict(costmaps [n_patches,ph,pw,OC] in the patch dtype or None,
    device_ptr or None, refs [n_points,C], src_obs [n_points])"""
    ctx = ctx or _capi.default_context()
    interp = interp or _capi.default_interp()
    cfg = cfg or _capi.default_costmap_config()
    d = problem.desc()
    n_points = len(problem.xyz)
    if refs is None:
        refs = np.zeros((n_points, problem.channels))
    else:
        refs = np.ascontiguousarray(refs, np.float64)
    src = np.full(n_points, -1, np.int64)
    oc = 3 if cfg.as_gradientfield else 1
    n_patches = d.n_patches if d.n_patches else d.n_obs
    out = np.zeros((n_patches, problem.ph, problem.pw, oc), problem.patch_np_dtype) if to_host else None
    dptr = C.c_void_p()
    s = _capi.make_summary(0)
    _capi.check(ctx.lib.pxr_costmaps_compute(ctx.handle, C.byref(d), C.byref(interp), C.byref(cfg), _p(refs), _p(src),
                                             _p(out) if to_host else None, C.byref(dptr) if to_device else None,
                                             C.byref(s)))
    return {"costmaps": out, "device_ptr": dptr.value if to_device else None, "refs": refs, "src_obs": src,
            "summary": _capi.summary_to_dict(s)}


def synth_patches_device(n_patches, ps, channels, uv0, field_id, seed=0, noise=0.01, ctx=None):
    """Device-side synthetic patch slab (fp16). Returns the device pointer (int)."""
    ctx = ctx or _capi.default_context()
    uv0 = np.ascontiguousarray(uv0, np.float64); field_id = np.ascontiguousarray(field_id, np.int64)
    out = C.c_void_p()
    _capi.check(ctx.lib.pxr_synth_patches_device(ctx.handle, C.byref(out), C.c_int64(n_patches), int(ps), int(channels),
                                                 _p(uv0), _p(field_id), C.c_uint64(seed), C.c_double(noise)))
    return out.value


def device_free(ptr, ctx=None):
    ctx = ctx or _capi.default_context()
    _capi.check(ctx.lib.pxr_device_free(ctx.handle, C.c_void_p(ptr)))


def memcpy_d2h(host_array, dev_ptr, nbytes, ctx=None):
    ctx = ctx or _capi.default_context()
    _capi.check(ctx.lib.pxr_memcpy_d2h(ctx.handle, _p(host_array), C.c_void_p(dev_ptr), C.c_size_t(nbytes)))


def ka_run(problem, interp=None, options=None, ctx=None):
    """pxr_ka_run == FeatureMetricKeypointOptimizer.run: refines problem.keypoints IN PLACE."""
    ctx = ctx or _capi.default_context()
    interp = interp or _capi.default_interp()
    options = options or _capi.default_ka_options()
    d = problem.desc()
    s = _capi.make_summary(0)
    _capi.check(ctx.lib.pxr_ka_run(ctx.handle, C.byref(d), C.byref(interp), C.byref(options), C.byref(s)))
    return _capi.summary_to_dict(s)


def graph_labels(node_image, edge_src, edge_dst, edge_sim):
    """compute_track_labels / compute_score_labels / compute_root_labels on flat arrays (host, bit-exact)."""
    lib = _capi.load_lib()
    node_image = np.ascontiguousarray(node_image, np.int32)
    es = np.ascontiguousarray(edge_src, np.int64); ed = np.ascontiguousarray(edge_dst, np.int64)
    sim = np.ascontiguousarray(edge_sim, np.float64)
    n = len(node_image)
    tl = np.zeros(n, np.int64); sc = np.zeros(n); rt = np.zeros(n, np.uint8)
    _capi.check(lib.pxr_graph_track_labels(C.c_int64(n), _p(node_image), C.c_int64(len(es)), _p(es), _p(ed), _p(sim), _p(tl)))
    _capi.check(lib.pxr_graph_score_labels(C.c_int64(n), C.c_int64(len(es)), _p(es), _p(ed), _p(sim), _p(tl), _p(sc)))
    _capi.check(lib.pxr_graph_root_labels(C.c_int64(n), _p(tl), _p(sc), _p(rt)))
    return tl, sc, rt


def ka_problem_labels(track_labels, max_per_problem=50):
    lib = _capi.load_lib()
    tl = np.ascontiguousarray(track_labels, np.int64)
    out = np.zeros(len(tl), np.int32); nb = C.c_int32()
    _capi.check(lib.pxr_ka_problem_labels(C.c_int64(len(tl)), _p(tl), int(max_per_problem), _p(out), C.byref(nb)))
    return out, nb.value


def ka_shard_plan(problem_weight, world):
    """pxr_shard_ka_problems: rank of every packed KA problem (whole problems per rank, no collective needed)"""
    lib = _capi.load_lib()
    w = np.ascontiguousarray(problem_weight, np.int64)
    out = np.zeros(len(w), np.int32)
    _capi.check(lib.pxr_shard_ka_problems(C.c_int32(len(w)), _p(w), int(world), _p(out)))
    return out


def ba_estimate_device_bytes(problem, options=None):
    """pxr_ba_estimate_device_bytes: what the solve will take on the device (host computation) ->
    dict(patches, state, reduced_system, total) in bytes"""
    lib = _capi.load_lib()
    d = problem.desc()
    out = [C.c_double(), C.c_double(), C.c_double()]
    _capi.check(lib.pxr_ba_estimate_device_bytes(C.byref(d), C.byref(options) if options is not None else None,
                                                 C.byref(out[0]), C.byref(out[1]), C.byref(out[2])))
    p, s, r = (o.value for o in out)
    return {"patches": p, "state": s, "reduced_system": r, "total": p + s + r}


class DeviceSlab:
    """A [N,H,W,C] array in device memory owned by libpxr (returned by pxr_extract_patches); exposes the CUDA array
    interface, so FeatureMap / DevicePatches take it like a torch tensor.  Freed with the object."""

    def __init__(self, ptr, shape, dtype, ctx):
        self.ptr, self.shape, self.dtype, self._ctx = int(ptr), tuple(int(v) for v in shape), np.dtype(dtype), ctx
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False),
                                         "version": 3, "strides": None}

    def __del__(self):
        try:
            if self.ptr:
                self._ctx.lib.pxr_device_free(self._ctx.handle, C.c_void_p(self.ptr))
                self.ptr = 0
        except Exception:      # interpreter shutdown
            pass


def extract_patches(dense, corners, patch_size, l2_normalize=True, out_dtype=np.float16, channels_first=True,
                    to_host=False, ctx=None):
    """pxr_extract_patches: windows of ONE dense map, gathered on the device.  `dense` is a host numpy array or anything
    with __cuda_array_interface__ (a contiguous torch.cuda tensor), [C,H,W] (or [H,W,C] with channels_first=False).
    -> DeviceSlab [n, ps, ps, C] (or a numpy array with to_host=True)"""
    ctx = ctx or _capi.default_context()
    cai = getattr(dense, "__cuda_array_interface__", None)
    if cai is not None:
        if cai.get("strides") is not None:
            raise ValueError("the dense map must be contiguous")
        shape, in_dtype, ptr, on_dev = tuple(cai["shape"]), np.dtype(cai["typestr"]), int(cai["data"][0]), 1
    else:
        dense = np.ascontiguousarray(dense)
        shape, in_dtype, ptr, on_dev = dense.shape, dense.dtype, dense.ctypes.data, 0
    if len(shape) == 4 and shape[0] == 1:
        shape = shape[1:]
    if len(shape) != 3 or in_dtype not in _capi.DTYPE_IDS or np.dtype(out_dtype) not in _capi.DTYPE_IDS:
        raise ValueError("a dense map is a [C,H,W] / [H,W,C] float16/32/64 array")
    ch, h, w = shape if channels_first else (shape[2], shape[0], shape[1])
    corners = np.ascontiguousarray(corners, np.int32).reshape(-1, 2)
    n = len(corners)
    out = np.zeros((n, patch_size, patch_size, ch), out_dtype) if to_host else None
    if n == 0:
        return out if to_host else DeviceSlab(0, (0, patch_size, patch_size, ch), out_dtype, ctx)
    dptr = C.c_void_p()
    _capi.check(ctx.lib.pxr_extract_patches(ctx.handle, C.c_void_p(ptr), on_dev, _capi.DTYPE_IDS[in_dtype], int(ch), int(h),
                                            int(w), int(bool(channels_first)), _p(corners), C.c_int64(n), int(patch_size),
                                            int(bool(l2_normalize)), _capi.DTYPE_IDS[np.dtype(out_dtype)],
                                            _p(out) if to_host else None, None if to_host else C.byref(dptr)))
    return out if to_host else DeviceSlab(dptr.value, (n, patch_size, patch_size, ch), out_dtype, ctx)


def interpolate_patches(patches, corners, scales, item_patch, xys, interp=None, upsampling_factor=1.0, ctx=None):
    """pxr_interpolate_descriptors: the (bicubic, optionally L2-normalised) descriptor of patch item_patch[i] of a host
    [N,H,W,C] patch array at IMAGE coordinates xys[i] -> [n_items, C] float64"""
    ctx = ctx or _capi.default_context()
    interp = interp or _capi.default_interp()
    patches = np.ascontiguousarray(patches)
    if patches.ndim != 4 or patches.dtype not in _capi.DTYPE_IDS:
        raise ValueError("patches must be a [N,H,W,C] float16/32/64 array")
    corners = np.ascontiguousarray(corners, np.int32).reshape(-1, 2)
    scales = np.ascontiguousarray(scales, np.float64).reshape(-1, 2)
    item_patch = np.ascontiguousarray(item_patch, np.int64).reshape(-1)
    xys = np.ascontiguousarray(xys, np.float64).reshape(-1, 2)
    if len(corners) != len(patches) or len(scales) != len(patches) or len(item_patch) != len(xys):
        raise ValueError("one corner and scale per patch, one patch index per query point")
    out = np.zeros((len(xys), patches.shape[3]))
    _capi.check(ctx.lib.pxr_interpolate_descriptors(
        ctx.handle, _p(patches), C.c_int64(len(patches)), _capi.DTYPE_IDS[patches.dtype], int(patches.shape[1]),
        int(patches.shape[2]), int(patches.shape[3]), _p(corners), _p(scales), C.c_double(upsampling_factor),
        C.c_int64(len(xys)), _p(item_patch), _p(xys), C.byref(interp), _p(out)))
    return out
