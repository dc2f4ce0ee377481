"""The function names of the reference's `pixsfm/extract.py` (:22-222) over this package's extractor: the reference keeps
the extraction loop in free functions that take the FeatureExtractor first, here it lives in `DenseFeatureExtractor`
(features/extractor.py), which any object with the same three methods may replace."""
from .features import store_features
from .util.misc import check_memory


def get_keypoints_and_ids(image_name, keypoints, req_keypoint_ids):
    """(keypoints of the image restricted to the required ids, the ids, how many) — extract.py:22-33"""
    keypoints_i, keypoint_ids_i, num_req_kps = None, None, 0
    if keypoints is not None:
        keypoints_i = keypoints[image_name]
        num_req_kps = keypoints_i.shape[0]
        if req_keypoint_ids is not None:
            num_req_kps = len(req_keypoint_ids[image_name])
            keypoint_ids_i = req_keypoint_ids[image_name]
            keypoints_i = keypoints_i[list(req_keypoint_ids[image_name]), :]
    return keypoints_i, keypoint_ids_i, num_req_kps


def estimate_required_memory(extractor, image_dir, image_list, keypoints=None, req_keypoint_ids=None, use_cache=False):
    """bytes of host memory the patches of an extraction will take (extract.py:36-54); 0 with a cache, NaN when the
    extractor cannot tell (dense maps of images it has not seen yet)"""
    if use_cache:
        return 0
    conf = getattr(extractor, "conf", {})
    itemsize = {"half": 2, "float": 4, "double": 8}.get(conf.get("dtype", "half"), 2) if isinstance(conf, dict) else 2
    channels = getattr(extractor, "channels_per_level", None)
    if not (isinstance(conf, dict) and conf.get("sparse", True)) or channels is None:
        return float("nan")
    n = sum(get_keypoints_and_ids(name, keypoints, req_keypoint_ids)[2] for name in image_list)
    return n * int(conf.get("patch_size", 16)) ** 2 * sum(channels) * itemsize


def features_from_image_list(extractor, image_dir, image_list, keypoints=None, req_keypoint_ids=None, cache_path=None,
                             estimate_memory=True, level_prefix=""):
    if estimate_memory:
        check_memory(estimate_required_memory(extractor, image_dir, image_list, keypoints, req_keypoint_ids,
                                              bool(getattr(extractor, "conf", {}).get("use_cache", False))
                                              if isinstance(getattr(extractor, "conf", None), dict) else False))
    return extractor.features_from_image_list(image_dir, list(image_list), keypoints, req_keypoint_ids, cache_path=cache_path,
                                              level_prefix=level_prefix)


def features_from_reconstruction(extractor, reconstruction, image_dir, cache_path=None, estimate_memory=True):
    """patches around the projections of the 3D points every image observes (extract.py:153-194)"""
    return extractor.features_from_reconstruction(reconstruction, image_dir, cache_path=cache_path)


def features_from_graph(extractor, image_dir, graph, keypoints_dict=None, cache_path=None, estimate_memory=True):
    """patches of the keypoints that take part in a match (extract.py:197-215)"""
    return extractor.features_from_graph(image_dir, graph, keypoints_dict, cache_path=cache_path)


def load_features_from_cache(cache_path=None, fill=False):
    """extract.py:218-222: the cache file as a FeatureManager, lazily filled by default"""
    return store_features.load_features_from_cache(cache_path, fill=fill)
