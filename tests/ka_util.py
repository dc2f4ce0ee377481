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


def make_query_ka_problem(n_queries=3, multi_ref_every=5, bound=4.0, **kw):
    """Query keypoint adjustment (localization/src/single_query_keypoint_optimizer.h): every image of a synthetic KA
    scene is one 'query' whose keypoints are pulled towards FIXED reference descriptors — here the (normalised)
    centre pixel of the next observation of the same track, every `multi_ref_every`-th keypoint against two of them."""
    sc = synthetic.make_ka_scene(**kw)
    n = len(sc["keypoints"])
    L = kw.get("track_len", 4)
    ps = sc["patches"].shape[1]
    centre = sc["patches"][:, ps // 2, ps // 2, :].astype(np.float64)
    centre /= np.linalg.norm(centre, axis=1, keepdims=True)
    refs, es, ed, ep = [], [], [], []
    img = sc["node_image"]
    order = np.argsort(img % n_queries, kind="stable")     # problem label = image id mod n_queries
    for i in order:
        tr, k = divmod(int(i), L)
        targets = [tr * L + (k + 1) % L]
        if i % multi_ref_every == 0:
            targets.append(tr * L + (k + 2) % L)
        for t in targets:
            es.append(int(i)); ed.append(len(refs)); ep.append(int(img[i] % n_queries)); refs.append(centre[t])
    prob = _capi.KAProblem(keypoints=sc["keypoints"], kp_const=np.zeros(n, np.uint8), edge_src=es, edge_dst=ed,
                           edge_weight=None, edge_problem=ep, n_problems=n_queries, patches=sc["patches"],
                           corner=sc["corner"], scale=sc["scale"], bound=bound, patches_are_sparse=True,
                           ref_desc=np.array(refs))
    return prob, sc
