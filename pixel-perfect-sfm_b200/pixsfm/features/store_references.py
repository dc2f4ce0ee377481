"""The reference cache file (reference features/store_references.py:14-58): one map {point3D_id: Reference} per feature
level, as the localization side consumes them (localization/main.py: `references[level]`).

    /                     attrs: n_levels
    /<level>/<point3D_id>/  descriptor [n_nodes, C] f64 | observations [n_obs, n_nodes, C] | costs [n_obs]
                            track_list_image_id [n] | track_list_point2D_idx [n]
                            attrs: source_image_id, source_point2D_idx

Through h5py when it is installed, util/h5lite.py otherwise."""
import numpy as np

from .._pixsfm._features import Reference
from ..util.hloc import _h5py


def _track_elements(track):
    """a pycolmap Track, a list of TrackElements or of (image_id, point2D_idx) pairs -> two integer lists"""
    if track is None:
        return [], []
    elements = getattr(track, "elements", track)
    pairs = [(e.image_id, e.point2D_idx) if hasattr(e, "image_id") else (e[0], e[1]) for e in elements]
    return [int(a) for a, _ in pairs], [int(b) for _, b in pairs]


def write_references_cache(path, list_map_id_refs):
    """Write list of maps of feature references to cache (one per level)"""
    with _h5py().File(str(path), "w") as f:
        f.attrs["n_levels"] = len(list_map_id_refs)
        for lvl, map_id_refs in enumerate(list_map_id_refs):
            lvl_grp = f.create_group(str(lvl))
            for pt_id, reference in map_id_refs.items():
                grp = lvl_grp.create_group(str(pt_id))
                descriptor = np.asarray(reference.descriptor, np.float64)
                grp.create_dataset("descriptor", data=descriptor)
                if len(reference.observations):
                    obs = np.asarray(reference.observations, np.float64).reshape((-1,) + descriptor.shape)
                else:
                    obs = np.zeros((0,) + descriptor.shape)
                grp.create_dataset("observations", data=obs)
                grp.create_dataset("costs", data=np.asarray(reference.costs, np.float64).reshape(-1))
                image_ids, point2D_idxs = _track_elements(reference.track)
                grp.create_dataset("track_list_image_id", data=np.asarray(image_ids, np.int64))
                grp.create_dataset("track_list_point2D_idx", data=np.asarray(point2D_idxs, np.int64))
                source = reference.source
                grp.attrs["source_image_id"] = int(source.image_id if hasattr(source, "image_id") else source[0])
                grp.attrs["source_point2D_idx"] = int(source.point2D_idx if hasattr(source, "point2D_idx") else source[1])


def load_references_from_cache(path):
    """Load list of maps of feature references from cache (one per level)"""
    f = _h5py().File(str(path), "r")
    try:
        out = []
        for lvl in range(int(np.asarray(f.attrs["n_levels"]).reshape(-1)[0])):
            refs = {}
            level = f[str(lvl)]
            for pt_id in level.keys():
                grp = level[pt_id]
                r = Reference((int(np.asarray(grp.attrs["source_image_id"]).reshape(-1)[0]),
                               int(np.asarray(grp.attrs["source_point2D_idx"]).reshape(-1)[0])),
                              np.array(grp["descriptor"], np.float64))
                obs = np.array(grp["observations"], np.float64)
                r.observations = [o for o in obs] if obs.size else []
                r.costs = [float(c) for c in np.asarray(grp["costs"]).reshape(-1)]
                r.track = list(zip((int(v) for v in np.asarray(grp["track_list_image_id"]).reshape(-1)),
                                   (int(v) for v in np.asarray(grp["track_list_point2D_idx"]).reshape(-1))))
                refs[int(pt_id)] = r
            out.append(refs)
        return out
    finally:
        if hasattr(f, "close"):
            f.close()
