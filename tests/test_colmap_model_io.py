"""COLMAP sparse-model files (text and binary) <-> the Reconstruction the bundle adjustment consumes
(pixsfm/util/colmap_model_io.py): exact round trips, the documented text layout, edge cases."""
import os
import struct

import numpy as np
import pytest

from pixsfm.util import colmap_model_io as mio
from pixsfm.util.colmap_types import INVALID_POINT3D, Reconstruction
from recon_util import make_reconstruction


def _same(a, b):
    assert sorted(a.cameras) == sorted(b.cameras) and sorted(a.images) == sorted(b.images)
    assert sorted(a.points3D) == sorted(b.points3D)
    for cid in a.cameras:
        ca, cb = a.cameras[cid], b.cameras[cid]
        assert (ca.model_id, ca.width, ca.height) == (cb.model_id, cb.width, cb.height)
        assert np.array_equal(ca.params, cb.params)
    for iid in a.images:
        ia, ib = a.images[iid], b.images[iid]
        assert ia.name == ib.name and ia.camera_id == ib.camera_id
        assert np.array_equal(ia.qvec, ib.qvec) and np.array_equal(ia.tvec, ib.tvec)
        assert len(ia.points2D) == len(ib.points2D)
        for pa, pb in zip(ia.points2D, ib.points2D):
            assert np.array_equal(pa.xy, pb.xy) and pa.point3D_id == pb.point3D_id
    for pid in a.points3D:
        pa, pb = a.points3D[pid], b.points3D[pid]
        assert np.array_equal(pa.xyz, pb.xyz)
        assert [(t.image_id, t.point2D_idx) for t in pa.track.elements] == [(t.image_id, t.point2D_idx) for t in pb.track.elements]


@pytest.mark.parametrize("ext", [".bin", ".txt"])
def test_round_trip_is_exact(tmp_path, ext):
    rec = make_reconstruction(n_cams=5, n_points=40, track_len=3, channels=16, seed=4)[0]
    some = rec.points3D[sorted(rec.points3D)[0]]
    some.color = np.array([10, 200, 30], np.uint8); some.error = 0.75
    mio.write_model(rec, tmp_path / "m", ext)
    assert sorted(os.listdir(tmp_path / "m")) == sorted(n + ext for n in ("cameras", "images", "points3D"))
    back = mio.read_model(tmp_path / "m")
    _same(rec, back)
    p = back.points3D[sorted(back.points3D)[0]]
    assert list(p.color) == [10, 200, 30] and p.error == 0.75
    assert back.num_observations() == rec.num_observations() and back.reg_image_ids() == rec.reg_image_ids()
    # the class methods pycolmap users call
    (rec.write if ext == ".bin" else rec.write_text)(str(tmp_path / "m2"))
    _same(rec, Reconstruction.read(tmp_path / "m2"))


def test_binary_layout_matches_the_published_format(tmp_path):
    rec = make_reconstruction(n_cams=3, n_points=10, track_len=2, channels=16, seed=1)[0]
    mio.write_model(rec, tmp_path, ".bin")
    raw = open(tmp_path / "cameras.bin", "rb").read()
    n, = struct.unpack_from("<Q", raw, 0)
    cid, model, w, h = struct.unpack_from("<iiQQ", raw, 8)
    cam = rec.cameras[sorted(rec.cameras)[0]]
    assert n == len(rec.cameras) and (cid, model, w, h) == (cam.camera_id, cam.model_id, cam.width, cam.height)
    assert struct.unpack_from("<%dd" % len(cam.params), raw, 8 + 24) == tuple(cam.params)
    raw = open(tmp_path / "points3D.bin", "rb").read()
    pid, = struct.unpack_from("<Q", raw, 8)
    p = rec.points3D[sorted(rec.points3D)[0]]
    assert pid == sorted(rec.points3D)[0] and struct.unpack_from("<3d", raw, 16) == tuple(p.xyz)
    assert struct.unpack_from("<Q", raw, 16 + 24 + 3 + 8)[0] == p.track.length()


def test_text_model_as_documented(tmp_path):
    (tmp_path / "cameras.txt").write_text("# Camera list with one line of data per camera:\n"
                                          "1 SIMPLE_PINHOLE 3072 2304 2559.81 1536 1152\n"
                                          "2 PINHOLE 3072 2304 2560.56 2560.56 1536 1152\n")
    (tmp_path / "images.txt").write_text("# Image list with two lines of data per image:\n"
                                         "1 0.851773 0.0165051 0.503764 -0.142941 -0.737434 1.02973 3.74354 1 P1180141.JPG\n"
                                         "2362.39 248.498 58396 1784.7 268.254 59027 1784.7 268.254 -1\n"
                                         "2 0.851773 0.0165051 0.503764 -0.142941 -0.737434 1.02973 3.74354 2 dir with space/P 2.JPG\n"
                                         "\n")
    (tmp_path / "points3D.txt").write_text("# 3D point list\n"
                                           "58396 3.68 1.6 10.2 215 180 140 0.42 1 0\n"
                                           "59027 1.0 2.0 3.0 1 2 3 1.5 1 1 2 5\n")
    rec = mio.read_model(tmp_path)
    assert rec.cameras[1].model_name == "SIMPLE_PINHOLE" and list(rec.cameras[2].params) == [2560.56, 2560.56, 1536, 1152]
    im = rec.images[1]
    assert im.name == "P1180141.JPG" and im.camera_id == 1 and np.allclose(im.qvec, [0.851773, 0.0165051, 0.503764, -0.142941])
    assert [p.point3D_id for p in im.points2D] == [58396, 59027, INVALID_POINT3D] and not im.points2D[2].has_point3D()
    assert rec.images[2].name == "dir with space/P 2.JPG" and rec.images[2].points2D == []
    assert [(t.image_id, t.point2D_idx) for t in rec.points3D[59027].track.elements] == [(1, 1), (2, 5)]
    assert list(rec.points3D[58396].color) == [215, 180, 140] and rec.points3D[58396].error == 0.42
    with pytest.raises(FileNotFoundError):
        mio.read_model(tmp_path / "nowhere")
