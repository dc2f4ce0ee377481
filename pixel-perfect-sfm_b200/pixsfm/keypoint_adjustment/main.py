"""pixsfm.keypoint_adjustment.main — the reference's KeypointAdjuster surface (pixsfm/keypoint_adjustment/main.py:13-279:
`create`, `refine_multilevel`, `refine`, `find_problem_labels`, `build_matching_graph`, same option names and values)
over the B200-backed optimizers."""
import numpy as np

from .. import base, defaults, logger
from .._pixsfm import _engine
from .._pixsfm import _keypoint_adjustment as ka
from ..util.conf import to_ctr
from ..util.refine import StrategyRefiner, optimizer_options


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


class KeypointAdjuster(StrategyRefiner):
    default_conf = defaults.keypoint_adjustment()
    optimizer_cls = None      # the `_keypoint_adjustment` optimizer a strategy drives

    def refine_multilevel(self, keypoints_dict, feature_manager, graph, track_labels=None, root_labels=None,
                          problem_setup=None):
        """one adjustment per feature level; `keypoints_dict` is refined in place"""
        if track_labels is None:
            track_labels = base.compute_track_labels(graph)
        if root_labels is None:
            root_labels = base.compute_root_labels(graph, track_labels, base.compute_score_labels(graph, track_labels))
        return self.per_level(feature_manager, lambda fset: self.refine(keypoints_dict, fset, graph, track_labels,
                                                                        root_labels, problem_setup=problem_setup))

    def refine(self, keypoints_dict, feature_set, graph, track_labels, root_labels, problem_setup=None):
        if problem_setup is None:        # the root of every track stays where it was detected
            problem_setup = ka.KeypointAdjustmentSetup()
            problem_setup.set_masked_nodes_constant(graph, root_labels)
        optimizer = self.optimizer_cls(optimizer_options(self.conf.optimizer, self.callbacks), problem_setup,
                                       to_ctr(self.conf.interpolation))
        run_args = (keypoints_dict, graph, track_labels, root_labels, feature_set)
        if self.conf.split_in_subproblems:
            run_args = (find_problem_labels(track_labels, self.conf.max_kps_per_problem)[0],) + run_args
        optimizer.run(*run_args)
        return {"summary": optimizer.summary()}


class FeatureMetricKeypointAdjuster(KeypointAdjuster, strategy="featuremetric"):
    """every matched pair of a track pulls on both of its keypoints (reference main.py:140-203)"""
    default_conf = defaults.keypoint_adjustment(root_regularize_weight=-1, weight_by_sim=True, root_edges_only=False)
    optimizer_cls = ka.FeatureMetricKeypointOptimizer


class TopologicalReferenceKeypointAdjuster(KeypointAdjuster, strategy="topological_reference"):
    """every keypoint of a track is pulled towards the node with the highest aggregated matching score (reference
    main.py:206-250): linear instead of quadratic in the track length"""
    optimizer_cls = ka.TopologicalReferenceKeypointOptimizer


def build_matching_graph(pairs, matches, scores=None):
    """matches of a COLMAP database / hloc match file -> base.Graph (reference main.py:262-271)"""
    logger.info("Building matching graph...")
    graph = base.Graph()
    for k, (name1, name2) in enumerate(pairs):
        graph.register_matches(name1, name2, matches[k], None if scores is None else scores[k])
    return graph


def extract_patchdata_from_graph(graph):
    """image name -> feature indices that take part in a match, i.e. the keypoints that need a patch
    (reference main.py:274-279)"""
    needed = {}
    for node in graph.nodes:
        needed.setdefault(graph.image_id_to_name[node.image_id], []).append(node.feature_idx)
    return needed
