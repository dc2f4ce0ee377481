"""What KeypointAdjuster and BundleAdjuster have in common: a strategy registry, the option dict handed to an optimizer,
and the loop over the feature levels (coarse to fine unless `level_indices` says otherwise, util/misc.py:19-23)."""
from .conf import merge, to_ctr


def optimizer_options(cfg, callbacks):
    """the optimizer's option dict with the solver callbacks injected (reference util/misc.py:30-36)"""
    options = to_ctr(cfg)
    options["solver"]["callbacks"] = callbacks
    return options


def level_order(level_indices, n_levels):
    return list(range(n_levels))[::-1] if level_indices in (None, "all") else list(level_indices)


class StrategyRefiner:
    """Base of the adjusters: `create(conf)` picks the subclass registered for conf.strategy; `per_level` runs one
    refinement per feature level and transposes the per-level result dicts into a dict of lists."""
    default_conf = {}
    callbacks = []
    _registry = None          # strategy name -> class; one dict per adjuster family (set on the family's base class)

    def __init_subclass__(cls, strategy=None, **kw):
        super().__init_subclass__(**kw)
        if StrategyRefiner in cls.__bases__:
            cls._registry = {}
        if strategy is not None:
            cls._registry[strategy] = cls

    @classmethod
    def create(cls, conf):
        name = conf["strategy"] if "strategy" in conf else cls.default_conf["strategy"]
        if name not in cls._registry:
            raise ValueError("strategy '%s' is not on the B200 path (%s)" % (name, ", ".join(sorted(cls._registry))))
        return cls._registry[name](conf)

    def __init__(self, conf):
        self.conf = merge(self.default_conf, conf)

    def per_level(self, feature_manager, refine_one):
        collected = {}
        for level in level_order(self.conf.level_indices, feature_manager.num_levels):
            for key, value in refine_one(feature_manager.fset(level)).items():
                collected.setdefault(key, []).append(value)
        return collected
