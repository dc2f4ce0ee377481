"""The names of the reference's `features/extract_patches.py` (:6-50).  There the gather runs in torch on the CNN's device
and the result is brought to numpy — "GPU->CPU remains main performance bottleneck" (:44); here the numpy form is
`cut_patches` (features/extractor.py) and the device form is libpxr's gather kernel (`pxr_extract_patches`), whose output
stays on the device."""
import numpy as np

from .extractor import cut_patches


def numpy_get_patch(arr, corner, ps):
    """one ps x ps window of an [H,W,C] (or [1,H,W,C]) map at corner (x, y)"""
    if arr.ndim == 4:
        return arr[0, corner[1]:corner[1] + ps, corner[0]:corner[0] + ps, :]
    if arr.ndim == 3:
        return arr[corner[1]:corner[1] + ps, corner[0]:corner[0] + ps, :]
    raise ValueError("a feature map is [H,W,C] or [1,H,W,C]")


def extract_patches_numpy(tensor, required_corners_np, ps):
    """[C,H,W] map (numpy, or a torch tensor) + corners [N,2] (x, y) -> [N, ps, ps, C] patches, rows = y"""
    if hasattr(tensor, "detach"):
        tensor = tensor.detach().cpu().numpy()
    chw = np.asarray(tensor)
    return cut_patches(np.moveaxis(chw, 0, -1), np.asarray(required_corners_np, np.int64).reshape(-1, 2), int(ps))


def extract_patches_device(tensor, required_corners_np, ps, l2_normalize=False, dtype=np.float16):
    """the device-side counterpart of the reference's extract_patches_torch: [C,H,W] map in host or device memory ->
    DevicePatches [N, ps, ps, C] that the optimizers consume without a round trip through the host"""
    from .._pixsfm import _engine
    return _engine.extract_patches(tensor, np.ascontiguousarray(required_corners_np, np.int32).reshape(-1, 2), int(ps),
                                   l2_normalize, dtype, True)
