"""CPU restatement (scipy) of the reference's adjacency construction -- TEST INFRASTRUCTURE ONLY.

  BaseData.create_adj_mat      beta_rec/data/base_data.py:337-360  (R[u, i] = 1 per interaction; A = [[0,R],[R^T,0]];
                                                                     norm = normalized_adj_single(A + I), mean = ...(A))
  normalized_adj_single        beta_rec/utils/common_util.py:24-41  (D^-1 adj with rowsum in float64, inf -> 0)
  sparse_mx_to_torch_sparse..  beta_rec/recommenders/lightgcn.py:15-23 (tocoo().astype(float32))
The Python loop over users is replaced by one coo_matrix construction (same matrix: duplicates collapse to 1);
checked against a literal transcription of the reference loop on a small graph in tests/test_adj_cpu.py.
"""
import numpy as np
import scipy.sparse as sp


def create_adj_mat(users, items, n_users, n_items):
    users, items = np.asarray(users), np.asarray(items)
    R = sp.coo_matrix((np.ones(len(users), dtype=np.float32), (users, items)), shape=(n_users, n_items)).tocsr()
    R.data[:] = 1.0  # R[u, item] = 1 however often the pair occurs
    n = n_users + n_items
    adj = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
    adj.resize((n, n))

    def normalized_adj_single(a):
        rowsum = np.array(a.sum(1))
        with np.errstate(divide="ignore"):
            d_inv = np.power(rowsum, -1).flatten()
        d_inv[np.isinf(d_inv)] = 0.0
        return sp.diags(d_inv).dot(a).tocoo()

    norm = normalized_adj_single(adj + sp.eye(n))
    mean = normalized_adj_single(adj.astype(np.float64))
    return adj.tocsr(), norm.tocsr(), mean.tocsr()


def to_coalesced_coo(mat):
    """(rows, cols, float32 values) in row-major sorted order without explicit zeros: what
    sparse_mx_to_torch_sparse_tensor(...).coalesce() holds."""
    m = mat.tocsr().astype(np.float32)
    m.sum_duplicates()
    m.sort_indices()
    m.eliminate_zeros()
    coo = m.tocoo()
    return coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data.astype(np.float32)
