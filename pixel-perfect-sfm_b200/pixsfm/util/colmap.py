"""Keypoints / matches of a COLMAP database <-> the containers the keypoint adjustment consumes
(reference: pixsfm/util/colmap.py:9-69; used by refine_colmap.py:105-112)."""
import numpy as np

from .database import COLMAPDatabase, blob_to_array, pair_id_to_image_ids


def read_image_id_to_name_from_db(database_path):
    db = COLMAPDatabase.connect(database_path)
    try:
        return db.image_id_to_name()
    finally:
        db.close()


def read_keypoints_from_db(database_path, as_cpp_map=True):
    """image name -> [N,2] float64 (x, y); scale / orientation columns are dropped.  `as_cpp_map` is accepted for
    signature compatibility: Map_NameKeypoints is a dict in this package."""
    db = COLMAPDatabase.connect(database_path)
    try:
        id2name = db.image_id_to_name()
        out = {}
        for image_id, rows, cols, data in db.execute("SELECT image_id, rows, cols, data FROM keypoints"):
            kp = blob_to_array(data, np.float32, (rows, cols)) if data is not None else np.zeros((0, 2), np.float32)
            out[id2name[image_id]] = np.ascontiguousarray(kp[:, :2], dtype=np.float64)
        return out
    finally:
        db.close()


def read_matches_from_db(database_path):
    """-> (pairs [(name1, name2)], matches [uint32 [M,2]], scores [float [M]] or None).  Scores are the cosine
    similarities of the matched (L2-normalised uint8) descriptors when the database holds descriptors."""
    db = COLMAPDatabase.connect(database_path)
    try:
        id2name = db.image_id_to_name()
        desc = {}
        for image_id, rows, cols, data in db.execute("SELECT image_id, rows, cols, data FROM descriptors"):
            d = blob_to_array(data, np.uint8, (-1, cols)).astype(np.float64)
            desc[image_id] = d / np.linalg.norm(d, axis=1, keepdims=True)
        scores = [] if desc else None
        pairs, matches = [], []
        for pair_id, data in db.execute("SELECT pair_id, data FROM matches"):
            if data is None:
                continue
            id1, id2 = pair_id_to_image_ids(pair_id)
            m = blob_to_array(data, np.uint32, (-1, 2))
            pairs.append((id2name[id1], id2name[id2]))
            matches.append(m)
            if scores is not None:
                scores.append(np.einsum("nd,nd->n", desc[id1][m[:, 0]], desc[id2][m[:, 1]]))
        return pairs, matches, scores
    finally:
        db.close()


def write_keypoints_to_db(database_path, keypoint_dict):
    """replaces the keypoints table with the (refined) keypoints, float32 [N,2]"""
    db = COLMAPDatabase.connect(database_path)
    try:
        db.execute("DELETE FROM keypoints")
        db.commit()
        name2id = {n: i for i, n in db.image_id_to_name().items()}
        for name, kp in keypoint_dict.items():
            db.add_keypoints(name2id[name], np.asarray(kp)[:, :2])
        db.commit()
    finally:
        db.close()
