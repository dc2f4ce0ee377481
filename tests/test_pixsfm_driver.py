"""PixSfM (pixsfm/refine_colmap.py): configuration handling and the parts of the driver that need no device."""
from pathlib import Path

import numpy as np
import pytest

from pixsfm.refine_colmap import PixSfM

YAML = """
interpolation:
  nodes: [[0.0, 0.0]]
  mode: BICUBIC
  l2_normalize: false
  ncc_normalize: false
dense_features:
  model: {name: s2dnet}
  patch_size: 16
mapping:
  interpolation: ${interpolation}
  KA:
    strategy: topological_reference
    interpolation: ${..interpolation}
    max_kps_per_problem: 20
  BA:
    strategy: costmaps
    interpolation: ${..interpolation}
    optimizer:
      refine_focal_length: false
"""


def test_defaults_are_the_references():
    s = PixSfM()
    assert s.conf.KA.strategy == "featuremetric" and s.conf.KA.max_kps_per_problem == 50
    assert s.conf.KA.optimizer.solver.parameter_tolerance == 1e-5 and s.conf.KA.optimizer.bound == 4.0
    assert s.conf.BA.strategy == "feature_reference" and s.conf.BA.optimizer.solver.use_inner_iterations is True
    assert s.conf.BA.optimizer.loss == {"name": "cauchy", "params": [0.25]}
    assert s.conf.KA.interpolation == s.conf.BA.interpolation == s.conf.interpolation
    assert type(s.keypoint_adjuster).__name__ == "FeatureMetricKeypointAdjuster"
    assert type(s.bundle_adjuster).__name__ == "FeatureReferenceBundleAdjuster"


def test_yaml_with_shared_interpolation_block(tmp_path):
    f = tmp_path / "conf.yaml"
    f.write_text(YAML)
    s = PixSfM(str(f))
    assert type(s.keypoint_adjuster).__name__ == "TopologicalReferenceKeypointAdjuster"
    assert type(s.bundle_adjuster).__name__ == "CostMapBundleAdjuster"
    assert s.conf.KA.max_kps_per_problem == 20 and s.conf.BA.optimizer.refine_focal_length is False
    assert s.conf.BA.optimizer.refine_extra_params is True                   # untouched defaults survive the merge
    for part in (s.conf.KA, s.conf.BA, s.conf):
        assert part.interpolation.l2_normalize is False and part.interpolation.mode == "BICUBIC"
    assert s.conf.dense_features.model.name == "s2dnet"


def test_own_interpolation_of_an_adjuster_wins():
    s = PixSfM({"interpolation": {"l2_normalize": False}, "BA": {"interpolation": {"l2_normalize": True, "mode": "BICUBIC",
                                                                                 "nodes": [[0.0, 0.0]], "ncc_normalize": False}}})
    assert s.conf.KA.interpolation.l2_normalize is False and s.conf.BA.interpolation.l2_normalize is True


def test_errors():
    with pytest.raises(ValueError, match="unknown configuration keys"):
        PixSfM({"KAA": {}})
    with pytest.raises(TypeError):
        PixSfM(3)
    with pytest.raises(ValueError, match="not on the B200 path"):
        PixSfM({"KA": {"strategy": "photometric"}})
    s = PixSfM()
    with pytest.raises(ValueError, match="no feature_manager given"):
        s.run_ba(object())
    with pytest.raises(NotImplementedError):
        s.triangulation()


def test_extractor_is_used_when_no_feature_manager_is_given():
    calls = []

    class Extractor:
        def features_from_reconstruction(self, reconstruction, image_dir, cache_path=None):
            calls.append(("rec", image_dir, cache_path))
            raise RuntimeError("stop here")

        def features_from_graph(self, image_dir, graph, keypoints, cache_path=None):
            calls.append(("graph", len(graph.nodes), sorted(keypoints)))
            raise RuntimeError("stop here")

    s = PixSfM(extractor=Extractor())
    with pytest.raises(RuntimeError):
        s.run_ba(object(), "imgs", cache_path="c.h5")
    with pytest.raises(RuntimeError):
        s.run_ka({"a": np.zeros((2, 2)), "b": np.zeros((2, 2))}, "imgs", [("a", "b")], ([np.array([[0, 1]], np.uint32)], None))
    assert calls == [("rec", "imgs", Path("c.h5")), ("graph", 2, ["a", "b"])]     # resolve_cache_path hands on a Path


def test_named_presets_and_reference_resolution(tmp_path):
    from pixsfm.configs import default_configs, parse_config_path
    assert {"default", "low_memory", "norefine"} <= set(default_configs) and parse_config_path(None) is None
    with pytest.raises(FileNotFoundError, match="Not in the default configs"):
        parse_config_path("no_such_preset")
    low = PixSfM("low_memory")
    assert type(low.keypoint_adjuster).__name__ == "TopologicalReferenceKeypointAdjuster"
    assert type(low.bundle_adjuster).__name__ == "CostMapBundleAdjuster"
    assert low.conf.KA.optimizer.bound == 2.0 and low.conf.KA.max_kps_per_problem == 1000
    assert low.conf.BA.optimizer.refine_extrinsics is False and low.conf.BA.max_tracks_per_problem == 100
    assert low.conf.dense_features.patch_size == 8 and low.conf.KA.optimizer.solver.max_num_iterations == 100
    assert PixSfM("norefine").conf.KA.apply is False and PixSfM("default").conf.BA.strategy == "feature_reference"
    # "${..name}" / "${name}" anywhere in the mapping block refer to the file's top-level blocks
    f = tmp_path / "c.yaml"
    f.write_text("""
dense_features: {patch_size: 10}
interpolation: {l2_normalize: false, mode: BICUBIC, nodes: [[0.0, 0.0]], ncc_normalize: false}
mapping:
  dense_features: ${..dense_features}
  interpolation: ${interpolation}
  KA: {interpolation: "${..interpolation}"}
  BA: {interpolation: "${..missing_block}"}
""")
    s = PixSfM(str(f))
    assert s.conf.dense_features.patch_size == 10
    assert s.conf.KA.interpolation.l2_normalize is False and s.conf.BA.interpolation.l2_normalize is False
