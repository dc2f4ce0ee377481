"""Mirror of `pixsfm._pixsfm._features` (pixsfm/features/bindings.cc:38-300): FeaturePatch / FeatureMap /
FeatureSet / FeatureView / FeatureManager and Reference.  Patches stay numpy references (no copy), as
featuremap.cc:9-45 does; the device upload happens when an optimizer runs.  HDF5 loading (featuremap.cc:138-267)
lives in features/store_features.py; `LazyFeatureMap` is what it returns for `fill=False`."""
import numpy as np

kDenseId = 1000000  # util/src/types.h:33


class DevicePatches:
    """A [N,H,W,C] patch array that already lives in device memory (the features come out of
    the CNN on the GPU; the reference moves them GPU -> numpy -> FeatureMap and flags that round trip as its
    bottleneck, features/extractor.py:152-236, extract_patches.py:41).  Wraps anything with the CUDA array
    interface (a contiguous torch.cuda tensor, a cupy array): the optimizers copy it device-to-device into their
    slab, nothing crosses PCIe.  Keeps the owner alive."""

    _TYPESTR = {"<f2": np.float16, "<f4": np.float32, "<f8": np.float64}

    def __init__(self, owner):
        cai = getattr(owner, "__cuda_array_interface__", None)
        if cai is None:
            raise ValueError("object does not expose __cuda_array_interface__")
        if cai.get("strides") is not None:
            raise ValueError("device patches must be C-contiguous")
        if cai["typestr"] not in self._TYPESTR:
            raise ValueError("device patches must be float16/float32/float64")
        self.owner = owner
        self.shape = tuple(int(v) for v in cai["shape"])
        self.ndim = len(self.shape)
        self.dtype = np.dtype(self._TYPESTR[cai["typestr"]])
        self.ptr = int(cai["data"][0])
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.flags = {"C_CONTIGUOUS": True}

    def __len__(self):
        return self.shape[0]


class FeaturePatch:
    def __init__(self, data, corner, scale, upsampling_factor=1.0):
        self.data = data  # [H,W,C] view
        self.corner = np.asarray(corner, np.int32)
        self.scale = np.asarray(scale, np.float64)
        self.upsampling_factor = float(upsampling_factor)

    @property
    def shape(self):
        return self.data.shape

    @property
    def height(self):
        return self.data.shape[0]

    @property
    def width(self):
        return self.data.shape[1]

    @property
    def channels(self):
        return self.data.shape[2]

    def to_pixel_coordinates(self, xy):  # featurepatch.h:250-255
        return (np.asarray(xy) * self.scale - 0.5 - self.corner) * self.upsampling_factor


class FeatureMap:
    """FeatureMap(patches[N,H,W,C] C-contiguous, point2D_ids, corners[N,2] (x,y) int32, metadata{"scale","is_sparse"})"""

    def __init__(self, patches, point2D_ids, corners, metadata):
        if hasattr(patches, "__cuda_array_interface__"):
            patches = DevicePatches(patches)        # stays on the device
        elif not isinstance(patches, DevicePatches):
            patches = np.asarray(patches)
        if patches.ndim != 4 or not patches.flags["C_CONTIGUOUS"]:
            raise ValueError("patches must be a C-contiguous [N,H,W,C] array")
        if patches.dtype not in (np.float16, np.float32, np.float64):
            raise ValueError("patches must be float16/float32/float64")
        self.patches = patches
        self.point2D_ids = [int(i) for i in point2D_ids]
        if len(self.point2D_ids) != patches.shape[0]:
            raise ValueError("number of point2D_ids and patches differ")
        self.corners = np.ascontiguousarray(corners, np.int32).reshape(-1, 2)
        self.scale = np.asarray(metadata["scale"], np.float64).reshape(2)
        self.is_sparse = bool(metadata.get("is_sparse", True))
        self._index = {pid: k for k, pid in enumerate(self.point2D_ids)}

    @property
    def shape(self):
        return self.patches.shape[1:]

    @property
    def channels(self):
        return self.patches.shape[3]

    @property
    def dtype(self):
        return self.patches.dtype

    def size(self):
        return self.patches.shape[0]

    def has_point2D(self, point2D_idx):
        # featuremap.h:113-119: a dense map answers for every keypoint with its single kDensePatchId patch
        if not self.is_sparse:
            return len(self._index) == 1 and kDenseId in self._index
        return int(point2D_idx) in self._index

    def local_index(self, point2D_idx):
        if not self.is_sparse:
            return self._index[kDenseId]            # GetFeaturePatch (featuremap.h:104-111)
        return self._index[int(point2D_idx)]

    def fpatch(self, point2D_idx):
        k = self.local_index(point2D_idx)
        if isinstance(self.patches, DevicePatches):
            raise ValueError("patch data is device-resident; host views are not available")
        return FeaturePatch(self.patches[k], self.corners[k], self.scale)


class LazyFeatureMap(FeatureMap):
    """A FeatureMap whose metadata is resident and whose patches come from the cache file on demand — the reference's
    `FeatureMap(h5_group, fill=false)` with `Load()` / `Unload()` / `Lock()` (featuremap.h:43-71, featuremap.cc:60-136).
    Granularity is the whole map (one image): `load()` reads all its patches through `reader()` and counts a user,
    `unload()` drops them when the last user is gone and the map is not locked.  Reading `patches` without a `load()`
    loads implicitly (and leaves the data in place until an `unload()`)."""

    def __init__(self, reader, n_patches, patch_shape, dtype, point2D_ids, corners, metadata):
        self._reader, self._data, self._users, self._locked = reader, None, 0, False
        self._shape4 = (int(n_patches),) + tuple(int(v) for v in patch_shape)
        self._dtype = np.dtype(dtype)
        if len(self._shape4) != 4:
            raise ValueError("patches must be [N,H,W,C]")
        if self._dtype not in (np.float16, np.float32, np.float64):
            raise ValueError("patches must be float16/float32/float64")
        self.point2D_ids = [int(i) for i in point2D_ids]
        if len(self.point2D_ids) != self._shape4[0]:
            raise ValueError("number of point2D_ids and patches differ")
        self.corners = np.ascontiguousarray(corners, np.int32).reshape(-1, 2)
        self.scale = np.asarray(metadata["scale"], np.float64).reshape(2)
        self.is_sparse = bool(metadata.get("is_sparse", True))
        self._index = {pid: k for k, pid in enumerate(self.point2D_ids)}

    # -- the reference's Load / Unload / Lock
    @property
    def is_loaded(self):
        return self._data is not None

    def load(self):
        if self._data is None:
            data = np.ascontiguousarray(self._reader())
            if data.shape != self._shape4 or data.dtype != self._dtype:
                raise ValueError("cache holds %s %s, metadata says %s %s" % (data.shape, data.dtype, self._shape4, self._dtype))
            self._data = data
        self._users += 1
        return self

    def unload(self):
        self._users = max(0, self._users - 1)
        if self._users == 0 and not self._locked:
            self._data = None

    def lock(self):
        """features stay resident from now on (FeatureMap::Lock)"""
        if self._data is None:
            self.load(); self._users -= 1
        self._locked = True

    @property
    def patches(self):
        if self._data is None:
            self.load(); self._users -= 1
        return self._data

    # -- what the base class derives from the array
    @property
    def shape(self):
        return self._shape4[1:]

    @property
    def channels(self):
        return self._shape4[3]

    @property
    def dtype(self):
        return self._dtype

    def size(self):
        return self._shape4[0]

    def fpatch(self, point2D_idx):
        k = self.local_index(point2D_idx)
        return FeaturePatch(self.patches[k], self.corners[k], self.scale)


class FeatureSet:
    def __init__(self, channels=None, dtype=None):
        self._channels, self._dtype, self._maps = channels, dtype, {}

    @property
    def channels(self):
        return self._channels

    @property
    def dtype(self):
        return self._dtype

    def emplace(self, image_name, fmap):
        if self._channels is None:
            self._channels = fmap.channels
        if fmap.channels != self._channels:
            raise ValueError("FeatureMap has %d channels, FeatureSet expects %d" % (fmap.channels, self._channels))
        if self._dtype is None:
            self._dtype = fmap.dtype
        self._maps[image_name] = fmap

    __setitem__ = emplace

    def has_fmap(self, image_name):
        return image_name in self._maps

    def fmap(self, image_name):
        return self._maps[image_name]

    __getitem__ = fmap

    def keys(self):
        return self._maps.keys()


class FeatureManager:
    def __init__(self, channels_per_level, dtype=np.float16):
        self.fsets = [FeatureSet(c, np.dtype(dtype)) for c in channels_per_level]

    @property
    def num_levels(self):
        return len(self.fsets)

    def fset(self, level_index):
        return self.fsets[level_index]


class FeatureView:
    """FeatureView(feature_set, reconstruction | graph): resolves image ids to names (featureview.cc:70-126)."""

    def __init__(self, feature_set, source, *_):
        self.fset = feature_set
        if hasattr(source, "image_id_to_name"):
            self._id_to_name = dict(source.image_id_to_name)
        else:
            self._id_to_name = {iid: im.name for iid, im in source.images.items()}
        # a view over a lazily filled set brings in the maps of ITS images and releases them when it goes away
        # (featureview.cc:70-126 loads the required patches in the constructor, the destructor unloads them)
        self._loaded = []
        for name in (dict.fromkeys(self._id_to_name.values()) if hasattr(feature_set, "has_fmap") else ()):
            if feature_set.has_fmap(name):
                fmap = feature_set.fmap(name)
                if isinstance(fmap, LazyFeatureMap):
                    fmap.load()
                    self._loaded.append(fmap)

    def close(self):
        for fmap in self._loaded:
            fmap.unload()
        self._loaded = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def channels(self):
        return self.fset.channels

    def image_name(self, image_id):
        return self._id_to_name[image_id]

    def has_feature_patch(self, image_id, point2D_idx):
        name = self._id_to_name.get(image_id)
        return name is not None and self.fset.has_fmap(name) and self.fset.fmap(name).has_point2D(point2D_idx)

    def get_feature_map(self, image_id):
        return self.fset.fmap(self._id_to_name[image_id])

    def get_feature_patch(self, image_id, point2D_idx):
        return self.get_feature_map(image_id).fpatch(point2D_idx)


class Reference:
    """features/src/references.h:32-65 (fields used on the named path)"""

    def __init__(self, source=(0, 0), descriptor=None):
        self.source = source  # (image_id, point2D_idx)
        self.descriptor = descriptor  # [n_nodes, C] float64
        self.track, self.observations, self.costs = None, [], []

    @property
    def source_image_id(self):
        return self.source[0]

    @property
    def source_point2D_idx(self):
        return self.source[1]

    def channels(self):
        return self.descriptor.shape[1]

    def num_nodes(self):
        return self.descriptor.shape[0]


class PatchSlab:
    """Device-upload plan for a set of FeatureMaps: blocks in first-use order + global patch indices."""

    def __init__(self):
        self.blocks, self.corners, self.scales, self._offset, self._n = [], [], [], {}, 0

    def index(self, image_name, fmap, point2D_idx):
        if image_name not in self._offset:
            self._offset[image_name] = self._n
            self.blocks.append(fmap.patches)
            self.corners.append(fmap.corners)
            self.scales.append(np.tile(fmap.scale, (fmap.size(), 1)))
            self._n += fmap.size()
        return self._offset[image_name] + fmap.local_index(point2D_idx)

    def arrays(self):
        return self.blocks, np.concatenate(self.corners), np.concatenate(self.scales)


class PatchInterpolator:
    """_features.PatchInterpolator(interpolation_config) (features/bindings.cc:276-292; dynamic_patch_interpolator.h:
    57-132, patch_interpolator.h:125-158): the descriptor of ONE FeaturePatch at a point, evaluated by libpxr.
      interpolate(patch, xy)        xy in image coordinates            -> [1, C]
      interpolate_nodes(patch, xy)  one row per interpolation node     -> [n_nodes, C]  (the device path has one node)
      interpolate_local(patch, uv)  uv = (column, row) inside the patch -> [1, C]"""

    def __init__(self, interpolation_config=None):
        from ._base import InterpolationConfig
        self.config = interpolation_config if isinstance(interpolation_config, InterpolationConfig) \
            else InterpolationConfig(interpolation_config or {})

    def _evaluate(self, patch, corner, scale, xy, upsampling_factor):
        from . import _capi, _engine
        self.config.validate_for_device()
        data = np.ascontiguousarray(patch.data)
        if data.ndim != 3:
            raise ValueError("a FeaturePatch holds an [H,W,C] array")
        ic = _capi.default_interp(self.config.l2_normalize, self.config.use_float_simd)
        return _engine.interpolate_patches(data[None], [corner], [scale], [0], [xy], ic, upsampling_factor)

    def interpolate(self, patch, xy):
        return self._evaluate(patch, patch.corner, patch.scale, np.asarray(xy, np.float64), patch.upsampling_factor)

    def interpolate_nodes(self, patch, xy):
        return self.interpolate(patch, xy)        # nodes = [[0, 0]]: the single node sits at xy itself

    def interpolate_local(self, patch, uv):
        # local coordinates are what ToPixelCoordinates produces: xy = uv + 0.5 with corner 0, scale 1 maps back onto uv
        return self._evaluate(patch, (0, 0), (1.0, 1.0), np.asarray(uv, np.float64) + 0.5, 1.0)
