"""pixsfm — B200-native drop-in for the featuremetric KA/BA hot path of cvg/pixel-perfect-sfm.

Mirrors the reference package name and the part of its surface that sits on the hot path; like the reference
(pixsfm/__init__.py:3-13) it owns a logger called "pixsfm" that prints INFO and above with a timestamp.  The compute
lives in csrc/ (CUDA, sm_100a) behind the C-ABI of include/pxr.h; there is no CPU fallback.
"""
import logging as _logging


def _make_logger(name="pixsfm", level=_logging.INFO):
    log = _logging.getLogger(name)
    log.setLevel(level)
    log.propagate = False            # do not print twice when the application configures the root logger
    if not log.handlers:
        out = _logging.StreamHandler()
        out.setLevel(level)
        out.setFormatter(_logging.Formatter("[%(asctime)s %(name)s %(levelname)s] %(message)s", "%Y/%m/%d %H:%M:%S"))
        log.addHandler(out)
    return log


logger = _make_logger()
