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
