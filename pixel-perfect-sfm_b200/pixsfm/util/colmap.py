"""Keypoints / matches of a COLMAP database <-> the containers the keypoint adjustment consumes.  Same four entry
points as the reference's pixsfm/util/colmap.py:9-69 (used by refine_colmap.py:105-112), on top of util/database.py."""
from contextlib import contextmanager

import numpy as np

from .database import COLMAPDatabase, blob_to_array, pair_id_to_image_ids


@contextmanager
def _database(path):
    db = COLMAPDatabase.connect(path)
    try:
        yield db
    finally:
        db.close()


def read_image_id_to_name_from_db(database_path):
    with _database(database_path) as db:
        return db.image_id_to_name()


def read_keypoints_from_db(database_path, as_cpp_map=True):
    """image name -> [N,2] float64 (x, y); scale / orientation columns are dropped.  `as_cpp_map` is accepted for
    signature compatibility: Map_NameKeypoints is a dict in this package."""
    with _database(database_path) as db:
        names = db.image_id_to_name()
        rows = db.execute("SELECT image_id, rows, cols, data FROM keypoints").fetchall()
    keypoints = {}
    for image_id, n, width, blob in rows:
        table = np.zeros((0, 2), np.float32) if blob is None else blob_to_array(blob, np.float32, (n, width))
        keypoints[names[image_id]] = np.ascontiguousarray(table[:, :2], dtype=np.float64)
    return keypoints


def _unit_descriptors(db):
    """image id -> L2-normalised descriptors as float64 (empty when the database stores none)"""
    unit = {}
    for image_id, _, width, blob in db.execute("SELECT image_id, rows, cols, data FROM descriptors"):
        raw = blob_to_array(blob, np.uint8, (-1, width)).astype(np.float64)
        unit[image_id] = raw / np.linalg.norm(raw, axis=1, keepdims=True)
    return unit


def read_matches_from_db(database_path):
    """-> (pairs [(name1, name2)], matches [uint32 [M,2]], scores [float [M]] or None).  A score is the cosine
    similarity of the two matched descriptors; None when the database holds no descriptors."""
    pairs, matches, scores = [], [], []
    with _database(database_path) as db:
        names = db.image_id_to_name()
        unit = _unit_descriptors(db)
        for pair_id, blob in db.execute("SELECT pair_id, data FROM matches"):
            if blob is None:          # pair without verified matches
                continue
            first, second = pair_id_to_image_ids(pair_id)
            idx = blob_to_array(blob, np.uint32, (-1, 2))
            pairs.append((names[first], names[second]))
            matches.append(idx)
            if unit:
                scores.append(np.sum(unit[first][idx[:, 0]] * unit[second][idx[:, 1]], axis=1))
    return pairs, matches, (scores if unit else None)


def write_keypoints_to_db(database_path, keypoint_dict):
    """replaces the keypoints table with the (refined) keypoints, float32 [N,2]"""
    with _database(database_path) as db:
        ids = {name: image_id for image_id, name in db.image_id_to_name().items()}
        db.execute("DELETE FROM keypoints")
        for name, xy in keypoint_dict.items():
            db.add_keypoints(ids[name], np.asarray(xy)[:, :2])
        db.commit()
