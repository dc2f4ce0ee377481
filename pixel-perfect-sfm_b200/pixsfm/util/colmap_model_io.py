"""COLMAP sparse-model files <-> util.colmap_types.Reconstruction, text (.txt) and binary (.bin).

The reference reads and writes models through pycolmap (`pycolmap.Reconstruction(path)` / `.write(path)`,
refine_colmap.py:117-131), which is not installable offline; this is an own implementation of the published COLMAP
model format (cameras / images / points3D; https://colmap.github.io/format.html):
  text    cameras.txt   CAMERA_ID MODEL WIDTH HEIGHT PARAMS[]
          images.txt    IMAGE_ID QW QX QY QZ TX TY TZ CAMERA_ID NAME  /  (X Y POINT3D_ID)*      (-1 = no 3D point)
          points3D.txt  POINT3D_ID X Y Z R G B ERROR (IMAGE_ID POINT2D_IDX)*
  binary  little endian, counts as uint64; camera: i32 id, i32 model, u64 w, u64 h, f64 params[];
          image: u32 id, f64 q[4], f64 t[3], u32 camera, name\\0, u64 n, (f64 x, f64 y, i64 point3D)*;
          point: u64 id, f64 xyz[3], u8 rgb[3], f64 error, u64 n, (u32 image, u32 point2D_idx)*
Colour and reprojection error are carried along (attributes `color`, `error`) so a read -> refine -> write cycle
keeps them."""
import os
import struct

import numpy as np

from .colmap_types import (Camera, Image, Point2D, Point3D, Reconstruction, Track, TrackElement, INVALID_POINT3D)
from .._pixsfm._capi import CAMERA_MODEL_IDS, CAMERA_NUM_PARAMS

_MODEL_NAMES = {v: k for k, v in CAMERA_MODEL_IDS.items()}


def _detect(path):
    for ext in (".bin", ".txt"):
        if all(os.path.isfile(os.path.join(path, n + ext)) for n in ("cameras", "images", "points3D")):
            return ext
    raise FileNotFoundError("no COLMAP model (cameras/images/points3D .bin or .txt) in %s" % path)


def _pid(v):
    return INVALID_POINT3D if v < 0 else int(v)


# ------------------------------------------------------------------------------------------ text
def _data_lines(fn):
    with open(fn) as f:
        for line in f:
            s = line.strip()
            if s and not s.startswith("#"):
                yield s


def _read_text(path, rec):
    for s in _data_lines(os.path.join(path, "cameras.txt")):
        e = s.split()
        rec.add_camera(Camera(int(e[0]), e[1], int(e[2]), int(e[3]), [float(x) for x in e[4:]]))
    # images.txt has exactly two lines per image; the second may be empty (no keypoints), so do not skip blanks there
    with open(os.path.join(path, "images.txt")) as f:
        lines = [l.rstrip("\n") for l in f if not l.startswith("#")]
    while lines and not lines[-1].strip():
        lines.pop()
    k = 0
    while k < len(lines):
        if not lines[k].strip():
            k += 1
            continue
        e = lines[k].split()
        pts = lines[k + 1].split() if k + 1 < len(lines) else []
        p2d = [Point2D((float(pts[i]), float(pts[i + 1])), _pid(int(pts[i + 2]))) for i in range(0, len(pts) - 2, 3)]
        rec.add_image(Image(int(e[0]), " ".join(e[9:]), int(e[8]), [float(x) for x in e[1:5]], [float(x) for x in e[5:8]], p2d))
        k += 2
    for s in _data_lines(os.path.join(path, "points3D.txt")):
        e = s.split()
        tr = Track([TrackElement(int(e[i]), int(e[i + 1])) for i in range(8, len(e) - 1, 2)])
        p = Point3D([float(x) for x in e[1:4]], tr)
        p.color = np.array([int(x) for x in e[4:7]], np.uint8)
        p.error = float(e[7])
        rec.points3D[int(e[0])] = p


def _write_text(path, rec):
    with open(os.path.join(path, "cameras.txt"), "w") as f:
        f.write("# Camera list with one line of data per camera:\n#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n")
        f.write("# Number of cameras: %d\n" % len(rec.cameras))
        for cid in sorted(rec.cameras):
            c = rec.cameras[cid]
            f.write("%d %s %d %d %s\n" % (cid, _MODEL_NAMES[c.model_id], c.width, c.height, " ".join(repr(float(x)) for x in c.params)))
    with open(os.path.join(path, "images.txt"), "w") as f:
        f.write("# Image list with two lines of data per image:\n#   IMAGE_ID, QW, QX, QY, QZ, TX, TY, TZ, CAMERA_ID, NAME\n"
                "#   POINTS2D[] as (X, Y, POINT3D_ID)\n# Number of images: %d\n" % len(rec.images))
        for iid in sorted(rec.images):
            im = rec.images[iid]
            f.write("%d %s %s %d %s\n" % (iid, " ".join(repr(float(x)) for x in im.qvec), " ".join(repr(float(x)) for x in im.tvec),
                                         im.camera_id, im.name))
            f.write(" ".join("%r %r %d" % (float(p.xy[0]), float(p.xy[1]), p.point3D_id if p.has_point3D() else -1)
                             for p in im.points2D) + "\n")
    with open(os.path.join(path, "points3D.txt"), "w") as f:
        f.write("# 3D point list with one line of data per point:\n#   POINT3D_ID, X, Y, Z, R, G, B, ERROR, TRACK[] as (IMAGE_ID, POINT2D_IDX)\n"
                "# Number of points: %d\n" % len(rec.points3D))
        for pid in sorted(rec.points3D):
            p = rec.points3D[pid]
            col = getattr(p, "color", np.zeros(3, np.uint8))
            f.write("%d %s %d %d %d %r %s\n" % (pid, " ".join(repr(float(x)) for x in p.xyz), col[0], col[1], col[2],
                                               float(getattr(p, "error", -1.0)),
                                               " ".join("%d %d" % (t.image_id, t.point2D_idx) for t in p.track.elements)))


# ------------------------------------------------------------------------------------------ binary
def _rd(f, fmt):
    fmt = "<" + fmt
    return struct.unpack(fmt, f.read(struct.calcsize(fmt)))


def _read_binary(path, rec):
    with open(os.path.join(path, "cameras.bin"), "rb") as f:
        for _ in range(_rd(f, "Q")[0]):
            cid, model, w, h = _rd(f, "iiQQ")
            rec.add_camera(Camera(cid, model, w, h, _rd(f, "%dd" % CAMERA_NUM_PARAMS[model])))
    with open(os.path.join(path, "images.bin"), "rb") as f:
        for _ in range(_rd(f, "Q")[0]):
            iid = _rd(f, "I")[0]
            q = _rd(f, "4d"); t = _rd(f, "3d")
            cam = _rd(f, "I")[0]
            name = bytearray()
            while True:
                ch = f.read(1)
                if ch in (b"\0", b""):
                    break
                name += ch
            n = _rd(f, "Q")[0]
            raw = np.frombuffer(f.read(24 * n), dtype=np.dtype([("x", "<f8"), ("y", "<f8"), ("p", "<i8")]))
            p2d = [Point2D((float(r["x"]), float(r["y"])), _pid(int(r["p"]))) for r in raw]
            rec.add_image(Image(iid, name.decode("utf-8"), cam, q, t, p2d))
    with open(os.path.join(path, "points3D.bin"), "rb") as f:
        for _ in range(_rd(f, "Q")[0]):
            pid = _rd(f, "Q")[0]
            xyz = _rd(f, "3d"); col = _rd(f, "3B"); err = _rd(f, "d")[0]
            n = _rd(f, "Q")[0]
            tr = np.frombuffer(f.read(8 * n), dtype="<u4").reshape(-1, 2)
            p = Point3D(xyz, Track([TrackElement(int(a), int(b)) for a, b in tr]))
            p.color = np.array(col, np.uint8); p.error = err
            rec.points3D[pid] = p


def _write_binary(path, rec):
    with open(os.path.join(path, "cameras.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(rec.cameras)))
        for cid in sorted(rec.cameras):
            c = rec.cameras[cid]
            f.write(struct.pack("<iiQQ", cid, c.model_id, c.width, c.height))
            f.write(np.asarray(c.params, "<f8").tobytes())
    with open(os.path.join(path, "images.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(rec.images)))
        for iid in sorted(rec.images):
            im = rec.images[iid]
            f.write(struct.pack("<I", iid))
            f.write(np.asarray(im.qvec, "<f8").tobytes()); f.write(np.asarray(im.tvec, "<f8").tobytes())
            f.write(struct.pack("<I", im.camera_id))
            f.write(im.name.encode("utf-8") + b"\0")
            f.write(struct.pack("<Q", len(im.points2D)))
            raw = np.zeros(len(im.points2D), dtype=np.dtype([("x", "<f8"), ("y", "<f8"), ("p", "<i8")]))
            for k, p in enumerate(im.points2D):
                raw[k] = (p.xy[0], p.xy[1], p.point3D_id if p.has_point3D() else -1)
            f.write(raw.tobytes())
    with open(os.path.join(path, "points3D.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(rec.points3D)))
        for pid in sorted(rec.points3D):
            p = rec.points3D[pid]
            col = getattr(p, "color", np.zeros(3, np.uint8))
            f.write(struct.pack("<Q", pid))
            f.write(np.asarray(p.xyz, "<f8").tobytes())
            f.write(struct.pack("<3Bd", int(col[0]), int(col[1]), int(col[2]), float(getattr(p, "error", -1.0))))
            f.write(struct.pack("<Q", len(p.track.elements)))
            f.write(np.array([[t.image_id, t.point2D_idx] for t in p.track.elements], "<u4").reshape(-1, 2).tobytes())


def read_model(path, ext=None):
    """-> Reconstruction read from a COLMAP model directory (binary preferred when both formats are present)"""
    path = str(path)
    ext = ext or _detect(path)
    rec = Reconstruction()
    (_read_binary if ext == ".bin" else _read_text)(path, rec)
    return rec


def write_model(rec, path, ext=".bin"):
    path = str(path)
    os.makedirs(path, exist_ok=True)
    (_write_binary if ext == ".bin" else _write_text)(path, rec)
