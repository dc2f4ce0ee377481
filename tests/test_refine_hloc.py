"""pixsfm.refine_hloc.PixSfM (reference pixsfm/refine_hloc.py:25-146): hloc keypoint / match files (HDF5) -> keypoint
adjustment -> refined keypoint file -> reconstruction (hloc / COLMAP, here a stub backend that writes a model) -> bundle
adjustment -> model on disk.  CPU: the plumbing with both refinements switched off.  GPU: the whole chain."""
import copy

import numpy as np
import pytest

from pixsfm import features
from pixsfm.refine_hloc import PixSfM, to_colmap_coordinates, to_hloc_coordinates
from pixsfm.util import hloc, synthetic
from pixsfm.util.colmap_types import Reconstruction
from recon_util import make_reconstruction


class _Backend:
    """stands in for hloc.reconstruction.main / hloc.triangulation.main: writes the given model where hloc would"""

    def __init__(self, rec):
        self.rec, self.calls = rec, []

    def reconstruction(self, model_path, image_dir, pairs_path, keypoints_path, matches_path, **kw):
        self.calls.append(("reconstruction", str(keypoints_path), kw))
        self.rec.write(str(model_path))

    def triangulation(self, model_path, reference_model_path, image_dir, pairs_path, keypoints_path, matches_path, **kw):
        self.calls.append(("triangulation", str(reference_model_path), str(keypoints_path)))
        self.rec.write(str(model_path))


def test_coordinate_conventions():
    kp = {"a": np.array([[1.0, 2.0]])}
    to_colmap_coordinates(kp); assert kp["a"].tolist() == [[1.5, 2.5]]
    to_hloc_coordinates(kp); assert kp["a"].tolist() == [[1.0, 2.0]]


def test_pipeline_plumbing_without_refinement(tmp_path):
    rec, fm, _, _ = make_reconstruction(n_cams=4, n_points=20, track_len=3, channels=16, seed=2)
    names = [rec.images[i].name for i in sorted(rec.images)]
    hloc.write_keypoints_hloc(tmp_path / "feats.h5", {n: np.zeros((3, 2)) for n in names})
    hloc.write_image_pairs(tmp_path / "pairs.txt", [(names[0], names[1])])
    back = _Backend(rec)
    sfm = PixSfM({"KA": {"apply": False}, "BA": {"apply": False}}, sfm_backend=back)
    out, data = sfm.reconstruction(tmp_path / "out", "imgs", tmp_path / "pairs.txt", tmp_path / "feats.h5", tmp_path / "matches.h5", my_arg=1)
    assert data["KA"] is None and data["BA"] is None
    assert back.calls == [("reconstruction", str(tmp_path / "feats.h5"), {"my_arg": 1})]     # unrefined keypoints go to hloc
    assert (tmp_path / "out" / "hloc").is_dir()
    written = Reconstruction.read(tmp_path / "out")
    assert sorted(written.images) == sorted(rec.images) and len(written.points3D) == len(rec.points3D)
    sfm.triangulation(tmp_path / "out2", tmp_path / "ref_model", "imgs", tmp_path / "pairs.txt", tmp_path / "feats.h5", tmp_path / "m.h5")
    assert back.calls[-1][0] == "triangulation" and back.calls[-1][1] == str(tmp_path / "ref_model")
    no_backend = PixSfM({"KA": {"apply": False}})
    if no_backend.sfm_backend is None:               # hloc is not installed here: the reference's error
        with pytest.raises(ValueError, match="Could not import hloc"):
            no_backend.run_reconstruction(tmp_path / "o3", "i", "p", "k", "m")


@pytest.mark.gpu
def test_keypoints_from_hloc_files_are_refined_then_the_model(tmp_path):
    from pixsfm import keypoint_adjustment as ka_pkg
    sc = synthetic.make_ka_scene(n_images=6, n_tracks=50, track_len=4, channels=128, seed=15, kp_sigma=1.0)
    names = ["db/im%d.jpg" % i for i in range(6)]
    kps = {names[i]: sc["keypoints"][sc["node_image"] == i] - 0.5 for i in range(6)}          # hloc convention: centre at integers
    hloc.write_keypoints_hloc(tmp_path / "feats.h5", kps)
    by_pair = {}
    for s, d in zip(sc["edge_src"], sc["edge_dst"]):
        a, b = int(sc["node_image"][s]), int(sc["node_image"][d])
        fa, fb = int(sc["node_feature"][s]), int(sc["node_feature"][d])
        if a > b:
            a, b, fa, fb = b, a, fb, fa
        by_pair.setdefault((a, b), []).append((fa, fb))
    pairs = [(names[a], names[b]) for a, b in by_pair]
    hloc.write_image_pairs(tmp_path / "pairs.txt", pairs)
    h5 = hloc._h5py()
    with h5.File(str(tmp_path / "matches.h5"), "w") as f:
        for (a, b), m in by_pair.items():
            m0 = np.full(len(kps[names[a]]), -1, np.int64); s0 = np.zeros(len(m0), np.float32)
            for fa, fb in m:
                m0[fa] = fb; s0[fa] = 0.9
            g = f.create_group(names[a].replace("/", "-") + "/" + names[b].replace("/", "-"))
            g.create_dataset("matches0", data=m0); g.create_dataset("matching_scores0", data=s0)
    fm = features.FeatureManager([128], np.float16)
    for i in range(6):
        m = np.where(sc["node_image"] == i)[0]
        fm.fset(0).emplace(names[i], features.FeatureMap(np.ascontiguousarray(sc["patches"][m]), sc["node_feature"][m].tolist(),
                                                          sc["corner"][m], {"scale": sc["scale"][m[0]], "is_sparse": True}))
    # the model "hloc" would build, with its own BA features
    rec, fm_ba, _, _ = make_reconstruction(n_cams=6, n_points=60, track_len=4, channels=128, seed=23)
    back = _Backend(rec)
    sfm = PixSfM({"KA": {"max_kps_per_problem": 20}, "BA": {"apply": False}}, sfm_backend=back)
    kp_ref, ka_data, _ = sfm.refine_keypoints(tmp_path / "refined.h5", tmp_path / "feats.h5", "imgs", tmp_path / "pairs.txt",
                                              tmp_path / "matches.h5", feature_manager=fm)
    assert ka_data["summary"][0].final_cost < 0.5 * ka_data["summary"][0].initial_cost
    stored = hloc.read_keypoints_hloc(tmp_path / "refined.h5")
    for n in names:
        assert np.array_equal(stored[n], kp_ref[n]) and np.abs(stored[n] - kps[n]).max() > 1e-3      # hloc convention on disk, moved
    # same solve as the direct adjuster on COLMAP-convention keypoints
    direct = {n: kps[n] + 0.5 for n in names}
    graph = ka_pkg.build_matching_graph(pairs, *hloc.read_matches_hloc(tmp_path / "matches.h5", pairs))
    ka_pkg.KeypointAdjuster.create({"max_kps_per_problem": 20}).refine_multilevel(direct, fm, graph)
    for n in names:
        assert np.abs(direct[n] - 0.5 - kp_ref[n]).max() < 1e-9
    # full run: KA (hloc files) -> "hloc" -> BA -> model on disk
    sfm2 = PixSfM({"KA": {"max_kps_per_problem": 20}, "BA": {"optimizer": {"solver": {"max_num_iterations": 6}}}}, sfm_backend=back)
    rec_before = copy.deepcopy(rec)
    # the two stages read different feature sets (keypoint patches / reconstruction patches): run them as `run` does,
    # handing each its own manager
    (tmp_path / "o").mkdir()
    sfm2.refine_keypoints(tmp_path / "o" / "refined_keypoints.h5", tmp_path / "feats.h5", "imgs", tmp_path / "pairs.txt",
                          tmp_path / "matches.h5", feature_manager=fm)
    model = sfm2.run_reconstruction(tmp_path / "o", "imgs", tmp_path / "pairs.txt", tmp_path / "o" / "refined_keypoints.h5", tmp_path / "matches.h5")
    loaded = Reconstruction.read(str(model))
    loaded, ba_data, _ = sfm2.run_ba(loaded, "imgs", feature_manager=fm_ba)
    assert ba_data["summary"][0].final_cost < ba_data["summary"][0].initial_cost
    moved = max(np.abs(loaded.points3D[p].xyz - rec_before.points3D[p].xyz).max() for p in loaded.points3D)
    assert moved > 1e-6 and back.calls[-1][0] == "reconstruction"
