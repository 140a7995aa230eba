"""GPU ranking evaluation (csrc/eval_kernels.cu through brs_rank_metrics) against the reference's own
known answers and against the eval oracle on seeded frames."""
import numpy as np
import pytest

from oracle import eval_oracle as E
from tests.test_eval_oracle import (EXPECTED, PERFECT, PRED_ITEMS, PRED_SCORES, TOL, TRUE_ITEMS, TRUE_RATINGS, USERS)

pytestmark = pytest.mark.gpu


def test_reference_known_answers_on_gpu():
    from beta_recsys_b200 import eval as G

    got = G.rank_metrics(USERS, TRUE_ITEMS, TRUE_RATINGS, USERS, PRED_ITEMS, PRED_SCORES, k=10)
    for m, want in EXPECTED.items():
        assert got[m] == pytest.approx(want, TOL), m
    got = G.rank_metrics(USERS, TRUE_ITEMS, TRUE_RATINGS, USERS, TRUE_ITEMS, TRUE_RATINGS, k=10)
    for m, want in PERFECT.items():
        assert got[m] == pytest.approx(want, 1e-12), m
    got = G.rank_metrics(USERS, TRUE_ITEMS, TRUE_RATINGS, USERS, [100] * 18, PRED_SCORES, k=10)
    assert all(v == 0.0 for v in got.values())


def _frame(rng, n_users, n_items, per_user, n_pos, ties):
    users = np.repeat(np.arange(n_users), per_user)
    items = np.concatenate([rng.choice(n_items, per_user, replace=False) for _ in range(n_users)])
    ratings = np.zeros(len(users), dtype=np.float32)
    for u in range(n_users):
        ratings[u * per_user + rng.choice(per_user, n_pos, replace=False)] = 1.0
    scores = rng.random(len(users)).astype(np.float32)
    if ties:
        scores = np.round(scores * 8) / 8  # many equal scores: exercises the row-order tie rule
    perm = rng.permutation(len(users))  # rows of a user are NOT contiguous in the frame
    return users[perm], items[perm], ratings[perm], scores[perm]


@pytest.mark.parametrize("ties", [False, True])
@pytest.mark.parametrize("k", [1, 5, 10, 33])
def test_evaluate_matches_oracle(k, ties):
    from beta_recsys_b200 import eval as G

    rng = np.random.default_rng(100 * k + ties)
    u, i, r, s = _frame(rng, 300, 2000, 101, 3, ties)
    want = E.rank_metrics(u, i, r, u, i, s, k)
    got = G.rank_metrics(u, i, r, u, i, s, k)
    for m in E.METRICS:
        assert got[m] == pytest.approx(want[m], rel=1e-9, abs=1e-12), (m, k, ties)


def test_evaluate_dataframe_signature_and_edge_cases():
    import pandas as pd

    from beta_recsys_b200 import eval as G

    rng = np.random.default_rng(7)
    u, i, r, s = _frame(rng, 50, 400, 40, 2, True)
    # a user without any relevant row, and a user id gap (ids need not be dense)
    u = np.concatenate([u, np.full(5, 77), np.full(3, 1000)])
    i = np.concatenate([i, np.arange(5), np.arange(3)])
    r = np.concatenate([r, np.zeros(5, np.float32), np.array([1, 0, 1], np.float32)])
    s = np.concatenate([s, rng.random(8).astype(np.float32)])
    df = pd.DataFrame({"col_user": u, "col_item": i, "col_rating": r})
    got = G.evaluate(df, s, ["ndcg", "precision", "recall", "map"], [5, 10])
    want = E.evaluate(u, i, r, s, ["ndcg", "precision", "recall", "map"], [5, 10])
    assert set(got) == set(want)
    for key in want:
        assert got[key] == pytest.approx(want[key], rel=1e-9, abs=1e-12), key
    with pytest.raises(NotImplementedError):
        G.evaluate(df, s, ["rmse"], 10)
    # empty prediction frame -> zeros, like the reference's `df_hit.shape[0] == 0` branch
    e = np.zeros(0, dtype=np.int64)
    assert all(v == 0.0 for v in G.rank_metrics(u, i, r, e, e, np.zeros(0, np.float32), 10).values())
    with pytest.raises(IndexError):
        G.rank_metrics(u, i, r, u, i, s, 10, n_user_ids=100)


def test_full_ranking_many_candidates():
    """All items as candidates for a few users (the full-ranking protocol): 20k rows per user."""
    from beta_recsys_b200 import eval as G

    rng = np.random.default_rng(11)
    u, i, r, s = _frame(rng, 8, 20000, 20000, 25, False)
    want = E.rank_metrics(u, i, r, u, i, s, 20)
    got = G.rank_metrics(u, i, r, u, i, s, 20)
    for m in E.METRICS:
        assert got[m] == pytest.approx(want[m], rel=1e-9, abs=1e-12), m
