"""CPU restatement of the negative sampler -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Semantics follow BaseData.instance_bpr_loader / instance_bce_loader (beta_rec/data/base_data.py:218-253,
182-216): one negative (BPR) or ``num_negative`` pairwise distinct negatives (BCE) per training row, uniform
over the items the row's user has NOT interacted with (``set(item_id_pool) - positive_items`` then
``random.sample``).  The reference draws from Python's global Mersenne Twister over a set's iteration order,
which is neither portable nor parallel; product and oracle share instead the counter-based stream
    candidate(row, t, attempt) = mix64(seed + row*C1 + t*C2 + attempt*C3) mod n_items
with rejection of the user's positives and of the row's earlier negatives, so the two are compared BIT FOR
BIT; the distribution (uniform over non-interacted items) is checked statistically in the tests.
"""
import numpy as np

C1, C2, C3 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xD1B54A32D192ED03), np.uint64(0x8CB92BA72F3D8DD7)
MAX_ATTEMPTS = 1 << 14


def mix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(30))
        x = x * np.uint64(0xBF58476D1CE4E5B9)
        x = x ^ (x >> np.uint64(27))
        x = x * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def sample_negatives(train_users, train_items, users, n_items, num_negative=1, seed=0):
    """negatives[r, t] for every row r of ``users``; positives are the (train_users, train_items) pairs."""
    pos = set(zip(np.asarray(train_users).tolist(), np.asarray(train_items).tolist()))
    users = np.asarray(users, dtype=np.int64)
    n = len(users)
    out = np.full((n, num_negative), -1, dtype=np.int64)
    rows = np.arange(n, dtype=np.uint64)
    seed = np.uint64(seed)
    with np.errstate(over="ignore"):
        for t in range(num_negative):
            todo = np.arange(n)
            for attempt in range(MAX_ATTEMPTS):
                if todo.size == 0:
                    break
                x = mix64(seed + rows[todo] * C1 + np.uint64(t) * C2 + np.uint64(attempt) * C3)
                cand = (x % np.uint64(n_items)).astype(np.int64)
                ok = np.fromiter(
                    ((int(users[r]), int(c)) not in pos and int(c) not in out[r, :t].tolist() for r, c in zip(todo, cand)),
                    dtype=bool, count=todo.size)
                out[todo[ok], t] = cand[ok]
                todo = todo[~ok]
            if todo.size:
                raise ValueError("a user has no item left to sample")
    return out


def bpr_triples(train_users, train_items, n_items, seed=0):
    """instance_bpr_loader's three tensors: (users, pos_items, neg_items), one negative per training row."""
    neg = sample_negatives(train_users, train_items, train_users, n_items, 1, seed)[:, 0]
    return np.asarray(train_users, dtype=np.int64), np.asarray(train_items, dtype=np.int64), neg


def bce_samples(train_users, train_items, train_ratings, n_items, num_negative, seed=0):
    """instance_bce_loader's three tensors: every training row followed by its num_negative negatives with
    rating 0 (base_data.py:203-210)."""
    neg = sample_negatives(train_users, train_items, train_users, n_items, num_negative, seed)
    u = np.repeat(np.asarray(train_users, dtype=np.int64), num_negative + 1)
    i = np.concatenate([np.asarray(train_items, dtype=np.int64)[:, None], neg], axis=1).reshape(-1)
    r = np.concatenate([np.asarray(train_ratings, dtype=np.float32)[:, None],
                        np.zeros((len(neg), num_negative), dtype=np.float32)], axis=1).reshape(-1)
    return u, i, r
