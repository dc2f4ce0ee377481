"""pixsfm.bundle_adjustment.main — same surface, key names and defaults as the reference's
pixsfm/bundle_adjustment/main.py:12-154; the bound classes are the B200-backed mirrors."""
from copy import deepcopy

from .. import features
from .._pixsfm import _bundle_adjustment as ba
from ..base import interpolation_default_conf, solver_default_conf
from ..util.conf import merge, to_ctr


def default_problem_setup(reconstruction):
    reg_image_ids = reconstruction.reg_image_ids()
    ba_setup = ba.BundleAdjustmentSetup()
    ba_setup.add_images(set(reg_image_ids))
    ba_setup.set_constant_pose(reg_image_ids[0])
    ba_setup.set_constant_tvec(reg_image_ids[1], [0])
    return ba_setup


def find_problem_labels(reconstruction, max_tracks_per_problem):
    problem_labels = [-1 for _ in range(max(reconstruction.point3D_ids()) + 1)]
    for p3D_id in reconstruction.point3D_ids():
        problem_labels[p3D_id] = int(p3D_id // max_tracks_per_problem)
    return problem_labels


def to_optim_ctr(cfg, callbacks):
    conf = to_ctr(cfg)
    conf["solver"]["callbacks"] = callbacks
    return conf


class BundleAdjuster:
    default_conf = {
        'apply': True,
        'interpolation': interpolation_default_conf,
        'level_indices': None,
        'max_tracks_per_problem': 10,
        'optimizer': {
            'loss': {'name': 'cauchy', 'params': [0.25]},
            'solver': {**solver_default_conf, 'use_inner_iterations': True},
            'print_summary': False,
            'refine_focal_length': True,
            'refine_principal_point': False,
            'refine_extra_params': True,
            'refine_extrinsics': True
        },
        'references': {
            'loss': {'name': 'cauchy', 'params': [0.25]},
            'iters': 100,
            'keep_observations': False,
            'compute_offsets3D': False,
            'num_threads': -1
        },
        'strategy': 'feature_reference'
    }
    callbacks = []

    @classmethod
    def create(cls, conf):
        strategy_to_solver = {"feature_reference": FeatureReferenceBundleAdjuster, "costmaps": CostMapBundleAdjuster}
        strategy = conf.get("strategy", cls.default_conf["strategy"])
        if strategy not in strategy_to_solver:
            raise ValueError("strategy '%s' is not on the B200 path (feature_reference, costmaps)" % strategy)
        return strategy_to_solver[strategy](conf)

    def refine(self, reconstruction, feature_set, problem_setup=None):
        raise NotImplementedError()

    def refine_multilevel(self, reconstruction, feature_manager, problem_setup=None):
        levels = self.conf.level_indices if self.conf.level_indices not in [None, "all"] else \
            list(reversed(range(feature_manager.num_levels)))
        outputs = {}
        for level_index in levels:
            out = self.refine(reconstruction, feature_manager.fset(level_index), problem_setup)
            for k, v in out.items():
                outputs.setdefault(k, []).append(v)
        return outputs


class FeatureReferenceBundleAdjuster(BundleAdjuster):
    """Featuremetric BA towards fixed per-point references (default method of the paper)."""
    default_conf = deepcopy(BundleAdjuster.default_conf)

    def __init__(self, conf):
        self.conf = merge(self.default_conf, conf)

    def refine(self, reconstruction, feature_set, problem_setup=None):
        if problem_setup is None:
            problem_setup = default_problem_setup(reconstruction)
        feature_view = features.FeatureView(feature_set, reconstruction)
        problem_labels = find_problem_labels(reconstruction, self.conf.max_tracks_per_problem)
        ref_extractor = ba.ReferenceExtractor(to_ctr(self.conf.references), to_ctr(self.conf.interpolation))
        references = ref_extractor.run(problem_labels, reconstruction, feature_set)
        solver = ba.FeatureReferenceBundleOptimizer(to_optim_ctr(self.conf.optimizer, self.callbacks), problem_setup,
                                                    to_ctr(self.conf.interpolation))
        solver.run(reconstruction, feature_view, references)
        return {"references": references, "summary": solver.summary()}


class CostMapBundleAdjuster(BundleAdjuster):
    """Cost-map BA (reference bundle_adjustment/main.py:218-286): cache, per observation, the robustified
    feature-metric cost towards the point's reference and its gradient as a 3-channel patch, then minimise the
    interpolated cost maps.  43x less patch memory than the 128-channel features."""
    default_conf = {
        **BundleAdjuster.default_conf,
        'costmaps': {
            'loss': {'name': 'trivial', 'params': []},
            'as_gradientfield': True,
            'compute_cross_derivative': False,
            'num_threads': -1
        },
    }

    def __init__(self, conf):
        self.conf = merge(self.default_conf, conf)

    def refine(self, reconstruction, feature_set, problem_setup=None):
        if problem_setup is None:
            problem_setup = default_problem_setup(reconstruction)
        problem_labels = find_problem_labels(reconstruction, self.conf.max_tracks_per_problem)
        interp_conf = to_ctr(self.conf.interpolation)
        ref_extractor = ba.ReferenceExtractor(to_ctr(self.conf.references), interp_conf)
        ce = ba.CostMapExtractor(to_ctr(self.conf.costmaps), interp_conf)
        costmap_fset, references = ce.run(problem_labels, reconstruction, feature_set, ref_extractor)
        interp_conf["l2_normalize"] = False   # "Make sure l2_normalize is set to false before optim!" (:262-263)
        costmap_view = features.FeatureView(costmap_fset, reconstruction)
        solver = ba.CostMapBundleOptimizer(to_optim_ctr(self.conf.optimizer, self.callbacks), problem_setup, interp_conf)
        solver.run(reconstruction, costmap_view)
        return {"costmaps": costmap_fset, "references": references, "summary": solver.summary()}
