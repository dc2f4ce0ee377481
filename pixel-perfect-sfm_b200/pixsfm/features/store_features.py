"""The dense-feature cache file (reference features/store_features.py:1-88 for the writer, features/src/featuremap.cc:
60-267 + featureset / featuremanager loaders for the reader; layout created by extract.py:98-128):

    /                         attrs: channels_per_level [int], dtype "half" | "float" | "double"
    /<level_prefix><level>/   one group per feature level
        <image name>/         (a name with "/" nests) attrs: format 1|2, is_sparse (int), further metadata (scale, patch_size…)
            format 2 "chunked":  patches [N,H,W,C] chunked (1,H,W,C) | keypoint_ids [N] | corners [N,2] | scales [N,2]
            format 1 "grouped":  attr shape [H,W,C]; one dataset per keypoint id with attrs corner, scale

Goes through h5py when it is installed and util/h5lite.py otherwise (classic HDF5 layout, see that module)."""
import numpy as np

from .._pixsfm._features import FeatureManager, FeatureMap, LazyFeatureMap, kDenseId
from ..util.hloc import _h5py

_DTYPES = {"half": np.float16, "float": np.float32, "double": np.float64}
_NAMES = {np.dtype(v): k for k, v in _DTYPES.items()}


def write_patch_cache(h5_parent, patch_id, data, corner, scale):
    dataset = h5_parent.create_dataset(str(patch_id), data=data)
    dataset.attrs["corner"] = corner
    dataset.attrs["scale"] = scale
    return dataset


def _write_metadata(h5_group, metadata):
    assert "is_sparse" in metadata
    for k, v in metadata.items():
        h5_group.attrs[k] = int(v) if k == "is_sparse" else v      # bools as ints (HighFive compatibility, store_features.py:34)


def write_featuremap_cache_grouped(h5_group, keypoint_ids, patches, corners, scales, metadata):
    h5_group.attrs["shape"] = list(patches.shape[1:])
    h5_group.attrs["format"] = 1
    _write_metadata(h5_group, metadata)
    for i, patch_id in enumerate(keypoint_ids):
        write_patch_cache(h5_group, patch_id, patches[i], corners[i], scales[i])
    return h5_group


def write_featuremap_cache_chunked(h5_group, keypoint_ids, patches, corners, scales, metadata):
    h5_group.attrs["format"] = 2
    _write_metadata(h5_group, metadata)
    chunks = [1, *patches.shape[1:]]
    if patches.shape[0] != len(keypoint_ids):          # a dense map stored once, read back as patches (store_features.py:62-64)
        chunks[1] = chunks[2] = metadata["patch_size"]
    h5_group.create_dataset("patches", data=patches, chunks=tuple(chunks))
    h5_group.create_dataset("keypoint_ids", data=np.asarray(keypoint_ids))
    h5_group.create_dataset("corners", data=np.asarray(corners))
    h5_group.create_dataset("scales", data=np.asarray(scales))
    return h5_group


def write_featuremap_cache(h5_group, keypoint_ids, patches, corners, scales, metadata, cache_format="chunked"):
    if cache_format == "grouped":
        return write_featuremap_cache_grouped(h5_group, keypoint_ids, patches, corners, scales, metadata)
    if cache_format == "chunked":
        return write_featuremap_cache_chunked(h5_group, keypoint_ids, patches, corners, scales, metadata)
    raise RuntimeError("Unknown cache_format %s to write." % cache_format)


def write_feature_manager_cache(path, feature_manager, cache_format="chunked", level_prefix=""):
    """the whole FeatureManager as extract.py:98-128 writes it level by level"""
    h5 = _h5py()
    with h5.File(str(path), "w") as f:
        f.attrs["channels_per_level"] = [int(feature_manager.fset(i).channels) for i in range(feature_manager.num_levels)]
        f.attrs["dtype"] = _NAMES[np.dtype(feature_manager.fset(0).dtype)]
        for level in range(feature_manager.num_levels):
            lg = f.create_group(level_prefix + str(level))
            fset = feature_manager.fset(level)
            for name in fset.keys():
                fmap = fset.fmap(name)
                scales = np.tile(np.asarray(fmap.scale, np.float64), (fmap.size(), 1))
                write_featuremap_cache(lg.create_group(name), fmap.point2D_ids, np.asarray(fmap.patches), fmap.corners, scales,
                                       {"scale": np.asarray(fmap.scale, np.float64), "is_sparse": fmap.is_sparse}, cache_format)


def _load_featuremap(group, point2D_ids=None):
    """FeatureMap::InitFromH5Group (featuremap.cc:60-267), both storage formats; `point2D_ids` restricts a sparse map"""
    fmt = int(group.attrs["format"])
    sparse = bool(int(group.attrs["is_sparse"]))
    if fmt == 2:
        ids = [int(v) for v in np.asarray(group["keypoint_ids"]).reshape(-1)]
        patches = np.asarray(group["patches"])
        corners = np.asarray(group["corners"]).reshape(-1, 2).astype(np.int32)
        scales = np.asarray(group["scales"]).reshape(-1, 2).astype(np.float64)
        if not sparse and len(ids) > 1:
            # "storing patch as dense but loading as sparse" (featuremap.cc:158-166): cut patch_size windows at the corners
            ps = int(group.attrs["patch_size"])
            patches = np.stack([patches[0, c[1]:c[1] + ps, c[0]:c[0] + ps] for c in corners])
            sparse = True
    elif fmt == 1:
        keys = group.keys() if hasattr(group, "keys") else list(group)
        ids = sorted(int(k) for k in keys) if sparse else [kDenseId]
        dsets = [group[str(k)] if sparse else group[list(keys)[0]] for k in ids]
        patches = np.stack([np.asarray(d) for d in dsets])
        corners = np.array([np.asarray(d.attrs["corner"]).reshape(2) for d in dsets], np.int32)
        scales = np.array([np.asarray(d.attrs["scale"]).reshape(2) for d in dsets], np.float64)
    else:
        raise RuntimeError("Unknown featuremap format.")
    if not sparse:
        ids = [kDenseId]
    if point2D_ids is not None and sparse:
        keep = [k for k, i in enumerate(ids) if i in set(int(p) for p in point2D_ids)]
        ids = [ids[k] for k in keep]; patches = patches[keep]; corners = corners[keep]; scales = scales[keep]
    scale = scales[0] if len(scales) else np.asarray(group.attrs["scale"], np.float64).reshape(2)
    return FeatureMap(np.ascontiguousarray(patches), ids, corners, {"scale": scale, "is_sparse": sparse})


def _lazy_featuremap(cache_path, group):
    """FeatureMap::InitFromH5Group with fill = false: metadata now, patches when a FeatureView needs the image.  Returns
    None for layouts whose metadata cannot be had without the data (they are read eagerly instead)."""
    fmt = int(group.attrs["format"])
    sparse = bool(int(group.attrs["is_sparse"]))
    if fmt != 2:
        return None                                   # "grouped": one dataset per keypoint, corner / scale on each
    ds = group["patches"]
    ids = [int(v) for v in np.asarray(group["keypoint_ids"]).reshape(-1)]
    if not sparse and len(ids) > 1:
        return None                                   # dense stored, sparse read: the windows are cut at load time
    corners = np.asarray(group["corners"]).reshape(-1, 2).astype(np.int32)
    scales = np.asarray(group["scales"]).reshape(-1, 2).astype(np.float64)
    scale = scales[0] if len(scales) else np.asarray(group.attrs["scale"], np.float64).reshape(2)
    if not sparse:
        ids = [kDenseId]
    name = group.name

    def reader():
        f = _h5py().File(str(cache_path), "r")
        try:
            return np.asarray(f[name]["patches"])
        finally:
            if hasattr(f, "close"):
                f.close()

    shape, dtype = tuple(ds.shape), np.dtype(ds.dtype)
    return LazyFeatureMap(reader, shape[0], shape[1:], dtype, ids, corners, {"scale": scale, "is_sparse": sparse})


def load_features_from_cache(cache_path, fill=True, level_prefix=""):
    """extract.py:218-222 / FeatureManager(path, fill, level_prefix): the cache file -> FeatureManager (numpy patches).
    `fill=False` (what the reference's low-memory configuration uses, extractor.py:49 `load_cache_on_init`) reads the
    metadata only: every map is a LazyFeatureMap that a FeatureView loads for the images it covers and drops again."""
    h5 = _h5py()
    f = h5.File(str(cache_path), "r")
    try:
        channels = [int(c) for c in np.asarray(f.attrs["channels_per_level"]).reshape(-1)]
        dtype = f.attrs["dtype"]
        dtype = dtype.decode() if isinstance(dtype, bytes) else str(dtype)
        fm = FeatureManager(channels, _DTYPES[dtype])

        def collect(group, prefix, fset):
            for key in group.keys():
                node = group[key]
                name = prefix + key
                if hasattr(node, "keys") and "format" in node.attrs:
                    fmap = None if fill else _lazy_featuremap(cache_path, node)
                    fset.emplace(name, fmap if fmap is not None else _load_featuremap(node))
                elif hasattr(node, "keys"):
                    collect(node, name + "/", fset)          # an image name with a directory part

        for level in range(len(channels)):
            collect(f[level_prefix + str(level)], "", fm.fset(level))
        return fm
    finally:
        if hasattr(f, "close"):
            f.close()
