"""The eval oracle against the known answers of the reference's own tests
(/root/reference/tests/test_evaluation.py: fixtures :31-141, expectations :270-420)."""
import numpy as np
import pytest

from oracle import eval_oracle as E

TOL = 0.0001  # the reference's own tolerance (test_evaluation.py:28)

# rating_true / rating_pred / rating_nohit fixtures of the reference, verbatim values
USERS = [1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 1, 1]
TRUE_ITEMS = [3, 1, 4, 5, 6, 7, 2, 5, 6, 8, 9, 10, 11, 12, 13, 14, 1, 2]
TRUE_RATINGS = [3, 5, 5, 3, 3, 1, 5, 5, 5, 4, 4, 3, 3, 3, 2, 1, 5, 4]
PRED_ITEMS = [12, 10, 3, 5, 11, 13, 4, 10, 7, 13, 1, 3, 5, 2, 11, 14, 3, 10]
PRED_SCORES = [12, 14, 13, 12, 11, 10, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 14, 13]
EXPECTED = {"ndcg": 0.38172, "map": 0.23613, "precision": 0.26666, "recall": 0.37777}
PERFECT = {"ndcg": 1.0, "map": 1.0, "precision": 0.6, "recall": 1.0}


def test_reference_known_answers():
    got = E.rank_metrics(USERS, TRUE_ITEMS, TRUE_RATINGS, USERS, PRED_ITEMS, PRED_SCORES, k=10)
    for m, want in EXPECTED.items():
        assert got[m] == pytest.approx(want, TOL), m


def test_perfect_and_no_hit():
    got = E.rank_metrics(USERS, TRUE_ITEMS, TRUE_RATINGS, USERS, TRUE_ITEMS, TRUE_RATINGS, k=10)
    for m, want in PERFECT.items():
        assert got[m] == pytest.approx(want, 1e-12), m
    got = E.rank_metrics(USERS, TRUE_ITEMS, TRUE_RATINGS, USERS, [100] * 18, PRED_SCORES, k=10)
    assert all(v == 0.0 for v in got.values())


def test_single_user_normalisation():
    # test_evaluation.py:320-345: precision of a 3-item user at k = 3 is 1, at k = 10 is 0.3
    u, i, r = [1, 1, 1], [1, 2, 3], [5, 4, 3]
    assert E.rank_metrics(u, i, r, u, i, r, k=3)["precision"] == pytest.approx(1.0, 1e-12)
    assert E.rank_metrics(u, i, r, u, i, r, k=10)["precision"] == pytest.approx(0.3, 1e-12)


def test_ties_keep_row_order_and_negatives_are_dropped():
    # two equal scores: the earlier row ranks first (nlargest keep="first", rank(method="first"))
    u = [0, 0, 0, 0]
    i = [10, 11, 12, 13]
    r = [0, 1, 0, 0]  # one relevant item; rating 0 rows are not truth (evaluation.py:492)
    s = [0.5, 0.5, 0.5, 0.1]
    got = E.evaluate(u, i, r, s, ["ndcg", "recall"], [1, 2])
    assert got["recall@1"] == 0.0 and got["recall@2"] == 1.0
    assert got["ndcg@2"] == pytest.approx((1 / np.log1p(2)) / (1 / np.log1p(1)), 1e-12)
