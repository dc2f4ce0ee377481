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

_SUBMODULES = ("base", "features", "bundle_adjustment", "keypoint_adjustment", "extract", "localization", "util", "residuals",
               "configs", "refine_colmap", "refine_hloc")


def __getattr__(name):
    """`import pixsfm; pixsfm.features...` as with the reference, whose __init__ imports its subpackages (:19-22); here they
    load on first use"""
    if name in _SUBMODULES:
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))


def set_debug():
    """reference __init__.py:28-30: DEBUG records from the "pixsfm" logger (the C++ side reports through status codes and
    pxr_last_error, it has no log level)"""
    logger.setLevel(_logging.DEBUG)
    for handler in logger.handlers:
        handler.setLevel(_logging.DEBUG)
