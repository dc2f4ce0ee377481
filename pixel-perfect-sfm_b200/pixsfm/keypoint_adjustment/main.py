"""pixsfm.keypoint_adjustment.main — same surface and defaults as the reference's
pixsfm/keypoint_adjustment/main.py:13-203."""
from copy import deepcopy

import numpy as np

from .. import base, logger
from .._pixsfm import _engine
from .._pixsfm import _keypoint_adjustment as ka
from ..bundle_adjustment.main import to_optim_ctr
from ..util.conf import merge, to_ctr


def find_problem_labels(track_labels, max_per_problem, track_edge_counts=None):
    """First-fit-decreasing packing of tracks into problems (reference main.py:13-57); host C++ in libpxr.so."""
    if track_edge_counts is not None:
        raise ValueError("track_edge_counts is not supported")
    labels, n_bins = _engine.ka_problem_labels(np.asarray(track_labels, np.int64), int(max_per_problem))
    bins = np.bincount(labels, minlength=n_bins)
    if max_per_problem > -1 and np.sum(bins > max_per_problem) > 0:
        logger.warning("%d / %d problems have more than %d keypoints.\n         Maximum keypoints in a problem: %d",
                       int(np.sum(bins > max_per_problem)), n_bins, max_per_problem, int(bins.max()))
    return [int(v) for v in labels], [int(b) for b in bins]


class KeypointAdjuster:
    default_conf = {
        'strategy': 'featuremetric',
        'apply': True,
        'interpolation': base.interpolation_default_conf,
        'level_indices': None,
        'max_kps_per_problem': 50,
        'optimizer': {
            'loss': {'name': 'cauchy', 'params': [0.25]},
            'solver': {**base.solver_default_conf, 'parameter_tolerance': 1.0e-5, 'num_threads': 1},
            'print_summary': False,
            'bound': 4.0,
            'num_threads': -1
        },
        'split_in_subproblems': True
    }
    callbacks = []

    @classmethod
    def create(cls, conf):
        strategy_to_solver = {"featuremetric": FeatureMetricKeypointAdjuster,
                              "topological_reference": TopologicalReferenceKeypointAdjuster}
        strategy = conf["strategy"] if "strategy" in conf else cls.default_conf["strategy"]
        if strategy not in strategy_to_solver:
            raise ValueError("strategy '%s' is not on the B200 path" % strategy)
        return strategy_to_solver[strategy](conf)

    def refine_multilevel(self, keypoints_dict, feature_manager, graph, track_labels=None, root_labels=None,
                          problem_setup=None):
        if track_labels is None:
            track_labels = base.compute_track_labels(graph)
        if root_labels is None:
            score_labels = base.compute_score_labels(graph, track_labels)
            root_labels = base.compute_root_labels(graph, track_labels, score_labels)
        levels = self.conf.level_indices if self.conf.level_indices not in [None, "all"] else \
            list(reversed(range(feature_manager.num_levels)))
        outputs = {}
        for level_index in levels:
            out = self.refine(keypoints_dict, feature_manager.fset(level_index), graph, track_labels, root_labels,
                              problem_setup=problem_setup)
            for k, v in out.items():
                outputs.setdefault(k, []).append(v)
        return outputs


class FeatureMetricKeypointAdjuster(KeypointAdjuster):
    default_conf = deepcopy(KeypointAdjuster.default_conf)
    default_conf["optimizer"] = {**default_conf["optimizer"], "root_regularize_weight": -1, "weight_by_sim": True,
                                 "root_edges_only": False, "num_threads": -1}

    def __init__(self, conf):
        self.conf = merge(self.default_conf, conf)

    def refine(self, keypoints_dict, feature_set, graph, track_labels, root_labels, problem_setup=None):
        if problem_setup is None:
            problem_setup = ka.KeypointAdjustmentSetup()
            problem_setup.set_masked_nodes_constant(graph, root_labels)
        solver = ka.FeatureMetricKeypointOptimizer(to_optim_ctr(self.conf.optimizer, self.callbacks), problem_setup,
                                                   to_ctr(self.conf.interpolation))
        if self.conf.split_in_subproblems:
            problem_labels, _ = find_problem_labels(track_labels, self.conf.max_kps_per_problem)
            solver.run(problem_labels, keypoints_dict, graph, track_labels, root_labels, feature_set)
        else:
            solver.run(keypoints_dict, graph, track_labels, root_labels, feature_set)
        return {"summary": solver.summary()}


class TopologicalReferenceKeypointAdjuster(KeypointAdjuster):
    """Optimize all keypoints of a track towards the node with the highest aggregated matching score (reference
    keypoint_adjustment/main.py:206-250): linear instead of quadratic in the track length."""
    default_conf = deepcopy(KeypointAdjuster.default_conf)
    default_conf["optimizer"] = {**default_conf["optimizer"], "num_threads": -1}

    def __init__(self, conf):
        self.conf = merge(self.default_conf, conf)

    def refine(self, keypoints_dict, feature_set, graph, track_labels, root_labels, problem_setup=None):
        if problem_setup is None:
            problem_setup = ka.KeypointAdjustmentSetup()
            problem_setup.set_masked_nodes_constant(graph, root_labels)
        solver = ka.TopologicalReferenceKeypointOptimizer(to_optim_ctr(self.conf.optimizer, self.callbacks), problem_setup,
                                                          to_ctr(self.conf.interpolation))
        if self.conf.split_in_subproblems:
            problem_labels, _ = find_problem_labels(track_labels, self.conf.max_kps_per_problem)
            solver.run(problem_labels, keypoints_dict, graph, track_labels, root_labels, feature_set)
        else:
            solver.run(keypoints_dict, graph, track_labels, root_labels, feature_set)
        return {"summary": solver.summary()}


def build_matching_graph(pairs, matches, scores=None):
    """matches of a COLMAP database / hloc match file -> base.Graph (reference main.py:262-271)"""
    logger.info("Building matching graph...")
    graph = base.Graph()
    scores = scores if scores is not None else [None] * len(matches)
    for (name1, name2), m, s in zip(pairs, matches, scores):
        graph.register_matches(name1, name2, m, s)
    return graph


def extract_patchdata_from_graph(graph):
    """image name -> feature indices that take part in a match, i.e. the keypoints that need a patch
    (reference main.py:274-279)"""
    needed = {}
    for node in graph.nodes:
        needed.setdefault(graph.image_id_to_name[node.image_id], []).append(node.feature_idx)
    return needed
