"""pixsfm — B200-native drop-in for the featuremetric KA/BA hot path of cvg/pixel-perfect-sfm.

Mirrors the reference package name and the part of its surface that sits on the hot path
(reference pixsfm/__init__.py:3-13 defines the same "pixsfm" logger).  The compute lives in
csrc/ (CUDA, sm_100a) behind the C-ABI of include/pxr.h; there is no CPU fallback.
"""
import logging

formatter = logging.Formatter(fmt="[%(asctime)s %(name)s %(levelname)s] %(message)s",
                              datefmt="%Y/%m/%d %H:%M:%S")
handler = logging.StreamHandler()
handler.setFormatter(formatter)
handler.setLevel(logging.INFO)

logger = logging.getLogger("pixsfm")
logger.setLevel(logging.INFO)
if not logger.handlers:
    logger.addHandler(handler)
logger.propagate = False
