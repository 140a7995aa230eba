"""The adjacency oracle against a literal transcription of the reference loop (base_data.py:337-360)."""
import numpy as np
import scipy.sparse as sp

from oracle import adj_oracle as A


def reference_loop(users, items, n_users, n_items):
    adj_mat = sp.dok_matrix((n_users + n_items, n_users + n_items), dtype=np.float32).tolil()
    R = sp.dok_matrix((n_users, n_items), dtype=np.float32)
    user_np, item_np = np.array(users), np.array(items)
    for u in range(n_users):
        for item in item_np[list(np.where(user_np == u)[0])]:
            R[u, item] = 1
    R = R.tolil()
    adj_mat[:n_users, n_users:] = R
    adj_mat[n_users:, :n_users] = R.T
    adj_mat = adj_mat.todok()

    def normalized_adj_single(adj):
        rowsum = np.array(adj.sum(1))
        with np.errstate(divide="ignore"):
            d_inv = np.power(rowsum, -1).flatten()
        d_inv[np.isinf(d_inv)] = 0.0
        return sp.diags(d_inv).dot(adj).tocoo()

    return normalized_adj_single(adj_mat + sp.eye(adj_mat.shape[0])).tocsr(), normalized_adj_single(adj_mat).tocsr()


def test_oracle_equals_reference_loop():
    rng = np.random.default_rng(0)
    n_users, n_items = 23, 17
    users, items = rng.integers(0, n_users, 120), rng.integers(0, n_items, 120)  # with duplicates
    users[users == 5] = 6  # user 5 has no interaction: its norm row is the self loop alone, its mean row empty
    _, norm, mean = A.create_adj_mat(users, items, n_users, n_items)
    rnorm, rmean = reference_loop(users, items, n_users, n_items)
    for got, want in ((norm, rnorm), (mean, rmean)):
        g, w = A.to_coalesced_coo(got), A.to_coalesced_coo(want)
        assert all(np.array_equal(x, y) for x, y in zip(g, w))
    r, c, v = A.to_coalesced_coo(norm)
    assert np.allclose(np.bincount(r, weights=v, minlength=n_users + n_items), 1.0, atol=1e-6)  # rows of D^-1(A+I) sum to 1
