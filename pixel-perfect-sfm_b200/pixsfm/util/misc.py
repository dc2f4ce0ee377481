"""The helper names of the reference's `pixsfm/util/misc.py` (:9-46), for code written against that module.  The
implementations live where this package uses them (util/conf.py, util/refine.py, refine_hloc.py)."""
from .. import logger
from .conf import to_ctr  # noqa: F401
from .refine import level_order as resolve_level_indices  # noqa: F401
from .refine import optimizer_options as to_optim_ctr  # noqa: F401


def free_memory():
    """bytes of host memory available (the reference asks its `_util.free_memory()`)"""
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except ImportError:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    return 0


def check_memory(req_memory, gap=2 ** 30):  # misc.py:9-16: warn when the estimate plus 1 GB exceeds what is free
    if req_memory != req_memory:
        logger.info("Invalid memory estimate. Continue.")
    elif req_memory + gap > free_memory():
        logger.warning("Warning: Required memory [%dMB] might exceed free memory [%dMB].", req_memory / 2 ** 20,
                       free_memory() / 2 ** 20)


def to_colmap_coordinates(keypoints):
    """hloc stores pixel centres at integers, COLMAP at .5 (misc.py:39-41); in place"""
    for name in keypoints.keys():
        keypoints[name] += 0.5


def to_hloc_coordinates(keypoints):
    """misc.py:44-46; in place"""
    for name in keypoints.keys():
        keypoints[name] -= 0.5
