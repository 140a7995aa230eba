"""Oracle-side checks of the negative sampler (no GPU): rejection semantics and the distribution."""
import numpy as np
import pytest

from oracle import sample_oracle as S


def test_negatives_avoid_positives_and_are_distinct_per_row():
    rng = np.random.default_rng(0)
    n_users, n_items = 40, 25
    tu, ti = rng.integers(0, n_users, 500), rng.integers(0, n_items, 500)
    neg = S.sample_negatives(tu, ti, tu, n_items, 5, seed=3)
    pos = set(zip(tu.tolist(), ti.tolist()))
    assert neg.min() >= 0 and neg.max() < n_items
    assert all((int(u), int(j)) not in pos for u, row in zip(tu, neg) for j in row)
    assert all(len(set(row)) == 5 for row in neg.tolist())
    assert np.array_equal(neg, S.sample_negatives(tu, ti, tu, n_items, 5, seed=3))  # reproducible
    assert not np.array_equal(neg, S.sample_negatives(tu, ti, tu, n_items, 5, seed=4))


def test_uniform_over_non_interacted_items():
    # one user who has interacted with items 0..9 of 20: negatives are uniform over 10..19
    tu, ti = np.zeros(10, dtype=np.int64), np.arange(10)
    users = np.zeros(20000, dtype=np.int64)
    neg = S.sample_negatives(tu, ti, users, 20, 1, seed=11)[:, 0]
    assert neg.min() >= 10
    counts = np.bincount(neg, minlength=20)[10:]
    assert abs(counts - 2000).max() < 5 * np.sqrt(2000 * 0.9)  # 5 sigma


def test_exhausted_user_raises_like_random_sample():
    tu, ti = np.zeros(4, dtype=np.int64), np.arange(4)
    with pytest.raises(ValueError):
        S.sample_negatives(tu, ti, tu, 4, 1, seed=0)


def test_bce_layout_matches_the_reference_loader():
    # base_data.py:203-210: row, then its negatives with rating 0
    tu, ti, tr = np.array([3, 1]), np.array([7, 2]), np.array([5.0, 4.0])
    u, i, r = S.bce_samples(tu, ti, tr, 10, 2, seed=1)
    assert u.tolist() == [3, 3, 3, 1, 1, 1]
    assert i[0] == 7 and i[3] == 2 and r.tolist() == [5.0, 0.0, 0.0, 4.0, 0.0, 0.0]
