"""Builds a KAProblem (flat IR) from a synthetic KA scene the way the reference's Python layer
does it: labels -> roots constant -> first-fit-decreasing problems -> intra-track edges grouped by problem
(keypoint_adjustment/main.py:168-203, topological_keypoint_optimizer.h:95-175)."""
import numpy as np

from pixsfm._pixsfm import _capi, _engine
from pixsfm.util import synthetic


def make_ka_problem(max_per_problem=50, bound=4.0, **kw):
    sc = synthetic.make_ka_scene(**kw)
    tl, scores, roots = _engine.graph_labels(sc["node_image"], sc["edge_src"], sc["edge_dst"], sc["edge_sim"])
    plabels, n_prob = _engine.ka_problem_labels(tl, max_per_problem)
    es, ed, sim = sc["edge_src"], sc["edge_dst"], sc["edge_sim"]
    intra = tl[es] == tl[ed]
    es, ed, sim = es[intra], ed[intra], sim[intra]
    eprob = plabels[es]
    order = np.argsort(eprob, kind="stable")
    prob = _capi.KAProblem(keypoints=sc["keypoints"], kp_const=roots, edge_src=es[order], edge_dst=ed[order],
                           edge_weight=sim[order], edge_problem=eprob[order], n_problems=n_prob,
                           patches=sc["patches"], corner=sc["corner"], scale=sc["scale"], bound=bound,
                           patches_are_sparse=True)
    return prob, sc, dict(track_labels=tl, roots=roots, problem_labels=plabels)
