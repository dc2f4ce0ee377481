"""Minimal stand-in for the OmegaConf usage of the reference's Python layer (omegaconf is not
installed offline): nested dicts with attribute access and a strict recursive merge.
Reference behaviour mirrored: `OmegaConf.merge(default_conf, conf)` (bundle_adjustment/main.py:115)."""
import copy


class Conf(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_conf(d):
    if isinstance(d, dict):
        return Conf({k: to_conf(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return [to_conf(v) for v in d]
    return d


def merge(default, override):
    out = to_conf(copy.deepcopy(default))
    if override is None:
        return out
    for k, v in override.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = merge(out[k], v)
        else:
            out[k] = to_conf(copy.deepcopy(v))
    return out


def to_ctr(cfg):
    """OmegaConf.to_container"""
    if isinstance(cfg, dict):
        return {k: to_ctr(v) for k, v in cfg.items()}
    if isinstance(cfg, list):
        return [to_ctr(v) for v in cfg]
    return cfg
