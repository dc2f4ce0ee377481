"""COLMAP camera models on the host (numpy): normalised image plane <-> pixels for the seven models the device path
supports, and world -> pixel for a posed camera.  Used where the Python layer has to PLACE things in an image (patch
corners around projected points, features/extractor.py); the optimisation itself projects on the device
(csrc/pxr_device.cuh::world_to_pixel).  Formulas: COLMAP src/base/camera_models.h (documented model definitions)."""
import numpy as np

from .._pixsfm._capi import CAMERA_MODEL_IDS, CAMERA_NUM_PARAMS


def _radial_tangential(u, v, k1, k2, p1, p2):
    r2 = u * u + v * v
    radial = k1 * r2 + k2 * r2 * r2
    return (u * radial + 2.0 * p1 * u * v + p2 * (r2 + 2.0 * u * u),
            v * radial + 2.0 * p2 * u * v + p1 * (r2 + 2.0 * v * v))


def normalized_to_image(model, params, uv):
    """uv [N,2] on the z = 1 plane -> pixels [N,2]"""
    model = CAMERA_MODEL_IDS[model] if isinstance(model, str) else int(model)
    p = np.asarray(params, np.float64)
    if len(p) < CAMERA_NUM_PARAMS[model]:
        raise ValueError("camera model %d takes %d parameters" % (model, CAMERA_NUM_PARAMS[model]))
    uv = np.asarray(uv, np.float64).reshape(-1, 2)
    u, v = uv[:, 0], uv[:, 1]
    single_focal = model in (0, 2, 3)
    fx, fy = (p[0], p[0]) if single_focal else (p[0], p[1])
    cx, cy = (p[1], p[2]) if single_focal else (p[2], p[3])
    du = dv = 0.0
    if model == 2:                                   # SIMPLE_RADIAL: f, cx, cy, k
        r2 = u * u + v * v
        du, dv = u * p[3] * r2, v * p[3] * r2
    elif model == 3:                                 # RADIAL: f, cx, cy, k1, k2
        r2 = u * u + v * v
        radial = p[3] * r2 + p[4] * r2 * r2
        du, dv = u * radial, v * radial
    elif model == 4:                                 # OPENCV: fx, fy, cx, cy, k1, k2, p1, p2
        du, dv = _radial_tangential(u, v, p[4], p[5], p[6], p[7])
    elif model == 5:                                 # OPENCV_FISHEYE: fx, fy, cx, cy, k1..k4
        r = np.sqrt(u * u + v * v)
        theta = np.arctan(r)
        t2 = theta * theta
        thetad = theta * (1.0 + t2 * (p[4] + t2 * (p[5] + t2 * (p[6] + t2 * p[7]))))
        with np.errstate(divide="ignore", invalid="ignore"):
            factor = np.where(r > np.finfo(np.float64).eps, thetad / r, 1.0)
        du, dv = u * factor - u, v * factor - v
    elif model == 6:                                 # FULL_OPENCV: fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, k5, k6
        r2 = u * u + v * v
        r4, r6 = r2 * r2, r2 * r2 * r2
        radial = (1.0 + p[4] * r2 + p[5] * r4 + p[8] * r6) / (1.0 + p[9] * r2 + p[10] * r4 + p[11] * r6)
        du = u * radial + 2.0 * p[6] * u * v + p[7] * (r2 + 2.0 * u * u) - u
        dv = v * radial + 2.0 * p[7] * u * v + p[6] * (r2 + 2.0 * v * v) - v
    return np.stack([fx * (u + du) + cx, fy * (v + dv) + cy], axis=1)


def rotation_matrix(qvec):
    """COLMAP quaternion (w, x, y, z), world -> camera"""
    w, x, y, z = np.asarray(qvec, np.float64) / np.linalg.norm(qvec)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def world_to_image(model, params, qvec, tvec, xyz):
    """pixels [N,2] of world points [N,3] seen by a camera at (qvec, tvec)"""
    pc = np.asarray(xyz, np.float64).reshape(-1, 3) @ rotation_matrix(qvec).T + np.asarray(tvec, np.float64)
    return normalized_to_image(model, params, pc[:, :2] / pc[:, 2:3])
