"""pixsfm/util/hloc.py: pair lists, the matches0 -> match-pair conversion, pair lookup, and the h5py gate."""
import numpy as np
import pytest

from pixsfm.util import hloc


def test_pair_list_round_trip(tmp_path):
    pairs = [("a.jpg", "b.jpg"), ("dir/c.jpg", "a.jpg")]
    hloc.write_image_pairs(tmp_path / "pairs.txt", pairs)
    assert (tmp_path / "pairs.txt").read_text() == "a.jpg b.jpg\ndir/c.jpg a.jpg"
    assert hloc.read_image_pairs(tmp_path / "pairs.txt") == [list(p) for p in pairs]
    (tmp_path / "p2.txt").write_text("x y\n\nz w\n")
    assert hloc.read_image_pairs(tmp_path / "p2.txt") == [["x", "y"], ["z", "w"]]


def test_matches0_conversion():
    m0 = np.array([-1, 4, -1, 0, 2])
    s0 = np.array([0.0, 0.9, 0.0, 0.5, 0.7])
    m, s = hloc.matches_from_hloc_arrays(m0, s0)
    assert m.dtype == np.uint64 and m.tolist() == [[1, 4], [3, 0], [4, 2]]
    assert s.dtype == np.float32 and np.allclose(s, [0.9, 0.5, 0.7])
    mr, _ = hloc.matches_from_hloc_arrays(m0, s0, reverse=True)
    assert mr.tolist() == [[4, 1], [0, 3], [2, 4]] and mr.flags["C_CONTIGUOUS"]
    m, s = hloc.matches_from_hloc_arrays(np.full(3, -1))
    assert m.shape == (0, 2) and s is None


def test_pair_lookup_like_hloc_find_pair():
    store = {"a.jpg/b.jpg", "x-c.jpg_a.jpg"}
    assert hloc._pair_key(store, "a.jpg", "b.jpg") == ("a.jpg/b.jpg", False)
    assert hloc._pair_key(store, "b.jpg", "a.jpg") == ("a.jpg/b.jpg", True)
    assert hloc._pair_key(store, "a.jpg", "x/c.jpg") == ("x-c.jpg_a.jpg", True)
    with pytest.raises(ValueError):
        hloc._pair_key(store, "a.jpg", "zzz.jpg")


def test_hdf5_keypoints_round_trip_with_or_without_h5py(tmp_path):
    """h5py when installed, util/h5lite.py otherwise (tests/test_h5lite.py covers that reader against real HDF5 files)"""
    kp = {"a.jpg": np.random.default_rng(0).uniform(0, 100, (5, 2))}
    hloc.write_keypoints_hloc(tmp_path / "kp.h5", kp)
    assert np.array_equal(hloc.read_keypoints_hloc(tmp_path / "kp.h5")["a.jpg"], kp["a.jpg"])
