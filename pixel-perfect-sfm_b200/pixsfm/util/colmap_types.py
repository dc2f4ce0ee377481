"""Duck-typed COLMAP containers (pycolmap is not installable offline).  Attribute names follow
pycolmap 0.4 as used by the reference (reconstruction.images[id].qvec / .tvec / .points2D[i],
.points3D[id].xyz / .track.elements, .cameras[id].params / .model_id, reg_image_ids(),
point3D_ids()).  Real pycolmap objects work with the adapters as well: only these attributes
are touched."""
import numpy as np

from .._pixsfm._capi import CAMERA_MODEL_IDS, CAMERA_NUM_PARAMS

INVALID_POINT3D = 2 ** 64 - 1


class Camera:
    def __init__(self, camera_id, model, width, height, params):
        self.camera_id = camera_id
        self.model_id = CAMERA_MODEL_IDS[model] if isinstance(model, str) else int(model)
        self.model_name = {v: k for k, v in CAMERA_MODEL_IDS.items()}[self.model_id]
        self.width, self.height = int(width), int(height)
        self.params = np.array(params, dtype=np.float64)
        assert len(self.params) == CAMERA_NUM_PARAMS[self.model_id]


class Point2D:
    def __init__(self, xy, point3D_id=INVALID_POINT3D):
        self.xy = np.array(xy, dtype=np.float64)
        self.point3D_id = point3D_id

    def has_point3D(self):
        return self.point3D_id != INVALID_POINT3D


class Image:
    def __init__(self, image_id, name, camera_id, qvec, tvec, points2D=()):
        self.image_id, self.name, self.camera_id = image_id, name, camera_id
        self.qvec = np.array(qvec, dtype=np.float64)
        self.tvec = np.array(tvec, dtype=np.float64)
        self.points2D = list(points2D)
        self.registered = True

    def num_points2D(self):
        return len(self.points2D)

    def normalize_qvec(self):
        self.qvec /= np.linalg.norm(self.qvec)


class TrackElement:
    def __init__(self, image_id, point2D_idx):
        self.image_id, self.point2D_idx = image_id, point2D_idx


class Track:
    def __init__(self, elements=()):
        self.elements = list(elements)

    def length(self):
        return len(self.elements)


class Point3D:
    def __init__(self, xyz, track=None):
        self.xyz = np.array(xyz, dtype=np.float64)
        self.track = track or Track()


class Reconstruction:
    def __init__(self):
        self.cameras, self.images, self.points3D = {}, {}, {}

    def add_camera(self, cam):
        self.cameras[cam.camera_id] = cam

    def add_image(self, img):
        self.images[img.image_id] = img

    def add_point3D(self, point3D_id, xyz):
        self.points3D[point3D_id] = Point3D(xyz)

    def add_observation(self, point3D_id, image_id, point2D_idx):
        self.points3D[point3D_id].track.elements.append(TrackElement(image_id, point2D_idx))
        self.images[image_id].points2D[point2D_idx].point3D_id = point3D_id

    def reg_image_ids(self):
        return sorted(i for i, im in self.images.items() if im.registered)

    def point3D_ids(self):
        return sorted(self.points3D.keys())

    def num_observations(self):
        return sum(p.track.length() for p in self.points3D.values())

    def as_arrays(self):
        """structure-of-arrays view of the topology (what pxr_recon_view takes): one pass over the Python objects.
        Works for pycolmap-like objects too (reconstruction_arrays)."""
        return reconstruction_arrays(self)

    # pycolmap.Reconstruction(path) / .write(path) / .write_text(path) (refine_colmap.py:117-131)
    @classmethod
    def read(cls, path):
        from .colmap_model_io import read_model
        return read_model(path)

    def write(self, path):
        from .colmap_model_io import write_model
        write_model(self, path, ".bin")

    def write_text(self, path):
        from .colmap_model_io import write_model
        write_model(self, path, ".txt")


def reconstruction_arrays(rec):
    """topology of a (duck-typed) reconstruction as flat arrays: the host-side input of pxr_problem_build"""
    image_ids = sorted(rec.images.keys())
    images = [rec.images[i] for i in image_ids]
    counts = np.fromiter((len(im.points2D) for im in images), np.int64, len(images))
    p2d_begin = np.zeros(len(images) + 1, np.int64)
    np.cumsum(counts, out=p2d_begin[1:])
    p2d_pid = np.full(int(p2d_begin[-1]), -1, np.int64)
    for k, im in enumerate(images):
        if counts[k]:
            # point3D_id is INVALID_POINT3D (2**64-1) where there is no 3D point: does not fit int64 -> -1
            p2d_pid[p2d_begin[k]:p2d_begin[k + 1]] = [p.point3D_id if p.point3D_id != INVALID_POINT3D else -1 for p in im.points2D]
    camera_ids = sorted(rec.cameras.keys())
    point_ids = sorted(rec.points3D.keys())
    tracks = [rec.points3D[p].track.elements for p in point_ids]
    tcounts = np.fromiter((len(t) for t in tracks), np.int64, len(tracks))
    track_begin = np.zeros(len(tracks) + 1, np.int64)
    np.cumsum(tcounts, out=track_begin[1:])
    n_el = int(track_begin[-1])
    return dict(image_id=np.array(image_ids, np.int64), image_camera_id=np.array([im.camera_id for im in images], np.int64),
                p2d_begin=p2d_begin, p2d_point3D_id=p2d_pid,
                camera_id=np.array(camera_ids, np.int64),
                camera_model=np.array([int(rec.cameras[c].model_id) for c in camera_ids], np.int32),
                point3D_id=np.array(point_ids, np.int64), track_begin=track_begin,
                track_image_id=np.fromiter((e.image_id for t in tracks for e in t), np.int64, n_el),
                track_point2D_idx=np.fromiter((e.point2D_idx for t in tracks for e in t), np.int64, n_el))


class ArrayReconstruction:
    """A reconstruction that IS arrays (no per-point Python objects): what a large model looks like when it is read
    straight into numpy, and what the bench's surface-level run uses.  Parameters live in `qvec [Ni,4]`, `tvec [Ni,3]`,
    `xyz [Np,3]`, `cam_params` (list of arrays) and are refined in place; `as_arrays()` hands the topology to
    pxr_problem_build without a pass over Python objects.  `images` / `cameras` / `points3D` give light read-only views
    for code that wants names or a single entry."""

    class _Image:
        def __init__(self, image_id, name, camera_id):
            self.image_id, self.name, self.camera_id = image_id, name, camera_id

    def __init__(self, image_ids, image_names, image_camera_ids, qvec, tvec, p2d_begin, p2d_point3D_id,
                 camera_ids, camera_models, cam_params, point3D_ids, xyz, track_begin, track_image_id, track_point2D_idx):
        self.image_id = np.ascontiguousarray(image_ids, np.int64)
        self.image_names = list(image_names)
        self.image_camera_id = np.ascontiguousarray(image_camera_ids, np.int64)
        self.qvec = np.ascontiguousarray(qvec, np.float64).reshape(-1, 4)
        self.tvec = np.ascontiguousarray(tvec, np.float64).reshape(-1, 3)
        self.p2d_begin = np.ascontiguousarray(p2d_begin, np.int64)
        self.p2d_point3D_id = np.ascontiguousarray(p2d_point3D_id, np.int64)
        self.camera_id = np.ascontiguousarray(camera_ids, np.int64)
        self.camera_model = np.ascontiguousarray(camera_models, np.int32)
        self.cam_params = [np.array(c, np.float64) for c in cam_params]
        self.point3D_id = np.ascontiguousarray(point3D_ids, np.int64)
        self.xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        self.track_begin = np.ascontiguousarray(track_begin, np.int64)
        self.track_image_id = np.ascontiguousarray(track_image_id, np.int64)
        self.track_point2D_idx = np.ascontiguousarray(track_point2D_idx, np.int64)
        for name in ("image_id", "camera_id", "point3D_id"):
            if np.any(np.diff(getattr(self, name)) <= 0):
                raise ValueError("%s must be strictly ascending" % name)
        self.images = {int(i): ArrayReconstruction._Image(int(i), n, int(c))
                       for i, n, c in zip(self.image_id, self.image_names, self.image_camera_id)}

    def as_arrays(self):
        return dict(image_id=self.image_id, image_camera_id=self.image_camera_id, p2d_begin=self.p2d_begin,
                    p2d_point3D_id=self.p2d_point3D_id, camera_id=self.camera_id, camera_model=self.camera_model,
                    point3D_id=self.point3D_id, track_begin=self.track_begin, track_image_id=self.track_image_id,
                    track_point2D_idx=self.track_point2D_idx)

    def num_observations(self):
        return int(self.track_begin[-1])
