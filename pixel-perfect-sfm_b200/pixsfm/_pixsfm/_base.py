"""Mirror of the reference's `pixsfm._pixsfm._base` (pixsfm/base/bindings.cc:29-157): the match
graph and the track / score / root labelling (base/src/graph.h:33-85, graph.cc:38-256).  The
labelling algorithms run in libpxr.so (host C++, bit-exact targets)."""
import numpy as np

from . import _engine


class InterpolatorType:
    BICUBIC = "BICUBIC"


class InterpolationConfig:
    """InterpolationConfig (base/src/interpolation.h:39-51); only the fields of the named path."""
    _fields = ("l2_normalize", "ncc_normalize", "nodes", "mode", "check_bounds", "use_float_simd")

    def __init__(self, conf=None, **kw):
        self.l2_normalize, self.ncc_normalize = True, False
        self.nodes, self.mode = [[0.0, 0.0]], "BICUBIC"
        self.check_bounds, self.use_float_simd = False, False
        self.mergedict(dict(conf or {}, **kw))

    def mergedict(self, d):
        for k, v in d.items():
            if k not in self._fields:
                raise ValueError("InterpolationConfig: unknown option '%s'" % k)  # strict keys (helpers.h:149-232)
            setattr(self, k, v)

    def todict(self):
        return {k: getattr(self, k) for k in self._fields}

    def validate_for_device(self):
        if str(self.mode).upper().split(".")[-1] != "BICUBIC":
            raise ValueError("only mode=BICUBIC is supported on the B200 path")
        if len(self.nodes) != 1 or self.ncc_normalize:
            raise ValueError("Unsupported dimensions (CHANNELS,N_NODES).")
        if self.check_bounds:
            raise ValueError("check_bounds=True is not supported on the B200 path")


class Match:
    __slots__ = ("node_idx", "sim")

    def __init__(self, node_idx, sim):
        self.node_idx, self.sim = node_idx, sim


class FeatureNode:
    __slots__ = ("image_id", "feature_idx", "node_idx", "out_matches")

    def __init__(self, image_id, feature_idx):
        self.image_id, self.feature_idx, self.node_idx, self.out_matches = image_id, feature_idx, -1, []


class Graph:
    """graph.cc:38-124.  Image ids inside the graph are insertion-order indices starting at 0."""

    def __init__(self):
        self.nodes = []
        self.image_name_to_id, self.image_id_to_name, self.node_map = {}, {}, {}

    def _image_id(self, name):
        if name not in self.image_name_to_id:
            iid = len(self.image_name_to_id)
            self.image_name_to_id[name] = iid
            self.image_id_to_name[iid] = name
        return self.image_name_to_id[name]

    def add_node(self, image, feature_idx):
        image_id = self._image_id(image) if isinstance(image, str) else image
        node = FeatureNode(image_id, int(feature_idx))
        node.node_idx = len(self.nodes)
        self.nodes.append(node)
        return node.node_idx

    def find_or_create_node(self, image_name, feature_idx):
        key = (self._image_id(image_name), int(feature_idx))
        if key not in self.node_map:
            self.node_map[key] = self.add_node(key[0], key[1])
        return self.nodes[self.node_map[key]]

    def add_edge(self, node1, node2, sim):
        node1.out_matches.append(Match(node2.node_idx, float(sim)))

    def register_matches(self, imname1, imname2, matches, similarities=None):
        matches = np.asarray(matches)
        for k in range(len(matches)):
            n1 = self.find_or_create_node(imname1, matches[k, 0])
            n2 = self.find_or_create_node(imname2, matches[k, 1])
            self.add_edge(n1, n2, 1.0 if similarities is None else similarities[k])

    def edges(self):
        return [(n.node_idx, m.node_idx, m.sim) for n in self.nodes for m in n.out_matches]

    def degrees(self):
        d = [0] * len(self.nodes)
        for n in self.nodes:
            d[n.node_idx] += len(n.out_matches)
            for m in n.out_matches:
                d[m.node_idx] += 1
        return d

    def scores(self):
        s = [0.0] * len(self.nodes)
        for n in self.nodes:
            for m in n.out_matches:
                s[m.node_idx] += m.sim; s[n.node_idx] += m.sim
        return s

    def flat(self):
        """(node_image[int32], es, ed, sim) in out_matches traversal order (the order graph.cc iterates in)"""
        node_image = np.array([n.image_id for n in self.nodes], np.int32)
        e = self.edges()
        es = np.array([x[0] for x in e], np.int64); ed = np.array([x[1] for x in e], np.int64)
        sim = np.array([x[2] for x in e], np.float64)
        return node_image, es, ed, sim


def compute_track_labels(graph):
    ni, es, ed, sim = graph.flat()
    return [int(v) for v in _engine.graph_labels(ni, es, ed, sim)[0]]


def compute_score_labels(graph, track_labels):
    import ctypes as C
    from . import _capi
    lib = _capi.load_lib()
    ni, es, ed, sim = graph.flat()
    tl = np.ascontiguousarray(track_labels, np.int64)
    sc = np.zeros(len(ni))
    _capi.check(lib.pxr_graph_score_labels(C.c_int64(len(ni)), C.c_int64(len(es)), es.ctypes.data_as(C.c_void_p),
                                           ed.ctypes.data_as(C.c_void_p), sim.ctypes.data_as(C.c_void_p),
                                           tl.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p)))
    return [float(v) for v in sc]


def compute_root_labels(graph, track_labels, score_labels):
    import ctypes as C
    from . import _capi
    lib = _capi.load_lib()
    tl = np.ascontiguousarray(track_labels, np.int64); sc = np.ascontiguousarray(score_labels, np.float64)
    rt = np.zeros(len(tl), np.uint8)
    _capi.check(lib.pxr_graph_root_labels(C.c_int64(len(tl)), tl.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                                          rt.ctypes.data_as(C.c_void_p)))
    return [bool(v) for v in rt]


def count_track_edges(graph, track_labels):
    n_tracks = len(set(track_labels))
    out = [0] * n_tracks
    for n in graph.nodes:
        for m in n.out_matches:
            if track_labels[n.node_idx] == track_labels[m.node_idx]:
                out[track_labels[n.node_idx]] += 1
    return out


def count_edges_AB(graph, track_labels, is_root):
    """per track: (edges that touch a root, edges between two non-roots) — graph.cc:258-281; the list has one entry per
    NODE index, as the reference allocates it, and only the entries of track labels are filled"""
    counts = [[0, 0] for _ in graph.nodes]
    for node in graph.nodes:
        for m in node.out_matches:
            t = track_labels[node.node_idx]
            if t == track_labels[m.node_idx]:
                counts[t][0 if (is_root[node.node_idx] or is_root[m.node_idx]) else 1] += 1
    return [tuple(c) for c in counts]
