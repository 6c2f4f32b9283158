"""CPU: the oracle of the helpers next to the path against hand-computed answers."""
import numpy as np

from oracle import utils_ref


def test_lfr_known_answers():
    x = np.arange(7 * 2, dtype=np.float64).reshape(7, 2)          # rows r = [2r, 2r+1]
    y = utils_ref.build_LFR_features(x, 4, 3)
    assert y.shape == (3, 8)                                       # ceil(7/3) rows of 4 frames
    assert y[0].tolist() == [0, 1, 2, 3, 4, 5, 6, 7]
    assert y[1].tolist() == [6, 7, 8, 9, 10, 11, 12, 13]
    assert y[2].tolist() == [12, 13, 12, 13, 12, 13, 12, 13]        # one real frame, then the LAST frame repeated
    assert np.array_equal(utils_ref.build_LFR_features(x, 1, 1), x)
    assert np.array_equal(utils_ref.build_LFR_features(x, 1, 2), x[::2])


def test_edit_distance_known_answers():
    assert utils_ref.levenshtein([1, 2, 3], [1, 2, 3]) == 0
    assert utils_ref.levenshtein([1, 2, 3], [1, 3]) == 1
    assert utils_ref.levenshtein([], [5, 6]) == 2
    assert utils_ref.levenshtein([7, 8, 9], []) == 3
    assert utils_ref.levenshtein([1, 2, 3, 4], [2, 3, 4, 5]) == 2   # kitten/sitting style shift
    d = utils_ref.edit_distance([[1, 2], [], [3]], [[1, 3, 2], [], []])
    assert d[0] == 1 / 3 and d[1] == 0.0 and np.isinf(d[2])
