"""COLMAP database I/O either side of the keypoint adjustment (pixsfm/util/colmap.py, database.py): blob encodings,
pair ids, the keypoint/match readers, and that a graph built from a database is the graph built from memory."""
import sqlite3

import numpy as np
import pytest

from pixsfm import base
from pixsfm.keypoint_adjustment import build_matching_graph, extract_patchdata_from_graph
from pixsfm.util import colmap as cio
from pixsfm.util.database import (COLMAPDatabase, MAX_IMAGE_ID, array_to_blob, blob_to_array, image_ids_to_pair_id,
                                  pair_id_to_image_ids)


def test_pair_id_is_colmaps():
    # COLMAP: pair_id = id_small * 2147483647 + id_large, independent of the argument order
    assert MAX_IMAGE_ID == 2147483647
    assert image_ids_to_pair_id(1, 2) == 2147483647 + 2 == image_ids_to_pair_id(2, 1)
    for a, b in [(1, 2), (7, 3), (2147483646, 5), (12, 12)]:
        assert pair_id_to_image_ids(image_ids_to_pair_id(a, b)) == (min(a, b), max(a, b))
    x = np.arange(12, dtype=np.float32).reshape(4, 3)
    assert np.array_equal(blob_to_array(array_to_blob(x), np.float32, (4, 3)), x)
    assert array_to_blob(x[:, ::2]) == np.ascontiguousarray(x[:, ::2]).tobytes()      # non-contiguous input


def _make_db(path, rng, n_images=4, n_kp=30, with_descriptors=True):
    db = COLMAPDatabase.connect(path)
    db.create_tables()
    cam = db.add_camera(2, 1000, 1000, [1200.0, 500.0, 500.0, 0.0])
    names, ids, kps, descs = [], [], {}, {}
    for i in range(n_images):
        name = "img_%02d.jpg" % i
        iid = db.add_image(name, cam)
        names.append(name); ids.append(iid)
        kp = np.concatenate([rng.uniform(0, 1000, (n_kp, 2)), rng.uniform(1, 3, (n_kp, 2))], 1)   # x, y, scale, orientation
        db.add_keypoints(iid, kp)
        kps[name] = kp[:, :2]
        if with_descriptors:
            d = rng.integers(0, 255, (n_kp, 128)).astype(np.uint8)
            db.add_descriptors(iid, d)
            descs[iid] = d
    pairs = {}
    for a in range(n_images):
        for b in range(a + 1, n_images):
            m = np.stack([rng.permutation(n_kp)[:12], rng.permutation(n_kp)[:12]], 1).astype(np.uint32)
            # give the images in both orders: the database stores the smaller id first
            if (a + b) % 2:
                db.add_matches(ids[b], ids[a], m[:, ::-1])
            else:
                db.add_matches(ids[a], ids[b], m)
            pairs[(names[a], names[b])] = m
    db.execute("INSERT INTO matches VALUES (?, ?, ?, ?)", (image_ids_to_pair_id(ids[0], ids[0]) + 999, 0, 2, None))
    db.commit(); db.close()
    return names, ids, kps, descs, pairs


def test_keypoints_and_matches_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    path = tmp_path / "database.db"
    names, ids, kps, descs, pairs = _make_db(path, rng)
    assert cio.read_image_id_to_name_from_db(path) == dict(zip(ids, names))
    got = cio.read_keypoints_from_db(path)
    assert isinstance(got, base.Map_NameKeypoints) and set(got) == set(names)
    for n in names:
        assert got[n].dtype == np.float64 and got[n].shape == (30, 2) and got[n].flags["C_CONTIGUOUS"]
        assert np.array_equal(got[n], kps[n].astype(np.float32).astype(np.float64))
    rp, rm, rs = cio.read_matches_from_db(path)
    assert len(rp) == len(pairs) == len(rm) == len(rs)          # the NULL-data row is skipped
    name2id = dict(zip(names, ids))
    for p, m, s in zip(rp, rm, rs):
        assert np.array_equal(m, pairs[p]) and m.dtype == np.uint32
        d1 = descs[name2id[p[0]]][m[:, 0]].astype(np.float64); d2 = descs[name2id[p[1]]][m[:, 1]].astype(np.float64)
        want = np.sum(d1 * d2, 1) / np.linalg.norm(d1, axis=1) / np.linalg.norm(d2, axis=1)
        assert np.allclose(s, want, rtol=0, atol=1e-12)
    # refined keypoints go back as float32 [N,2]
    refined = {n: got[n] + 0.25 for n in names}
    cio.write_keypoints_to_db(path, refined)
    back = cio.read_keypoints_from_db(path, as_cpp_map=False)
    for n in names:
        assert np.array_equal(back[n], refined[n].astype(np.float32).astype(np.float64))
    raw = sqlite3.connect(str(path)).execute("SELECT rows, cols FROM keypoints").fetchall()
    assert raw == [(30, 2)] * len(names)


def test_no_descriptors_means_no_scores(tmp_path):
    path = tmp_path / "d.db"
    _make_db(path, np.random.default_rng(1), n_images=3, with_descriptors=False)
    rp, rm, rs = cio.read_matches_from_db(path)
    assert rs is None and len(rp) == 3


def test_graph_from_database_equals_graph_from_memory(tmp_path):
    rng = np.random.default_rng(2)
    path = tmp_path / "g.db"
    names, ids, kps, descs, pairs = _make_db(path, rng)
    rp, rm, rs = cio.read_matches_from_db(path)
    g_db = build_matching_graph(rp, rm, rs)
    g_mem = base.Graph()
    for p, s in zip(rp, rs):
        g_mem.register_matches(p[0], p[1], pairs[p], s)
    assert g_db.edges() == g_mem.edges()
    assert [(n.image_id, n.feature_idx) for n in g_db.nodes] == [(n.image_id, n.feature_idx) for n in g_mem.nodes]
    tl = base.compute_track_labels(g_db)
    assert list(tl) == list(base.compute_track_labels(g_mem))
    needed = extract_patchdata_from_graph(g_db)
    assert set(needed) <= set(names)
    for n, idxs in needed.items():
        assert len(set(idxs)) == len(idxs) and max(idxs) < 30
    # unit similarities when the database has no descriptors
    g1 = build_matching_graph(rp, rm)
    assert all(e[2] == 1.0 for e in g1.edges())


def test_bad_shapes_raise(tmp_path):
    db = COLMAPDatabase.connect(tmp_path / "x.db")
    db.create_tables()
    with pytest.raises(ValueError):
        db.add_keypoints(1, np.zeros((3, 3)))
    with pytest.raises(ValueError):
        db.add_matches(1, 2, np.zeros((3, 3)))
    db.close()
