"""GPU adjacency builder (csrc/adj_kernels.cu) bit-exact against the oracle's matrix pushed through the
engine's host-side CSR conversion, and LightGCN trained on it."""
import numpy as np
import pytest
import torch

from oracle import adj_oracle as A

pytestmark = pytest.mark.gpu


def _graph(seed, n_users, n_items, e):
    rng = np.random.default_rng(seed)
    return rng.integers(0, n_users, e), rng.integers(0, n_items, e)


@pytest.mark.parametrize("mean", [False, True])
@pytest.mark.parametrize("shape", [(23, 17, 120), (700, 300, 20000), (5000, 9000, 100000)])
def test_adjacency_bit_exact(shape, mean):
    from beta_recsys_b200 import graph as G
    from beta_recsys_b200.engines.lightgcn import coo_to_csr

    n_users, n_items, e = shape
    users, items = _graph(e, n_users, n_items, e)
    users[users == 5] = 6
    adj = G.build_norm_adj(users, items, n_users, n_items, mean=mean)
    _, norm, mean_m = A.create_adj_mat(users, items, n_users, n_items)
    r, c, v = A.to_coalesced_coo(mean_m if mean else norm)
    want = coo_to_csr(r, c, v, n_users + n_items)
    got = {k: t.cpu().numpy() for k, t in adj.csr_tensors().items()}
    assert adj.nnz == want["nnz"]
    for k in ("row_ptr", "col", "val", "row_ptr_t", "col_t", "val_t", "edge_id_t"):
        assert np.array_equal(got[k], want[k]), k
    sp_t = adj.to_torch_sparse()
    assert np.array_equal(sp_t.indices().cpu().numpy(), np.vstack([r, c]))
    assert np.array_equal(sp_t.values().cpu().numpy(), v)


def test_adjacency_rejects_out_of_range():
    from beta_recsys_b200 import graph as G

    with pytest.raises(IndexError):
        G.build_norm_adj([0, 9], [1, 1], 5, 3)


def test_lightgcn_engine_accepts_gpu_adjacency():
    """Same training step whether the engine is given the reference's torch sparse tensor or the GPU-built
    adjacency (identical CSR arrays, so identical kernels on identical inputs)."""
    from beta_recsys_b200 import graph as G
    from beta_recsys_b200.engines import LightGCNEngine

    n_users, n_items = 60, 40
    users, items = _graph(3, n_users, n_items, 500)
    adj = G.build_norm_adj(users, items, n_users, n_items)

    def make(norm_adj):
        torch.manual_seed(0)
        cfg = {"model": dict(device_str="cuda:0", n_users=n_users, n_items=n_items, emb_dim=16, layer_size=[16, 16],
                             batch_size=64, optimizer="sgd", lr=0.05, regs=[1e-5], keep_pro=0.6, norm_adj=norm_adj),
               "system": {"run_dir": "/tmp/brs_test"}}
        return LightGCNEngine(cfg)

    e1, e2 = make(adj), make(adj.to_torch_sparse().cpu())
    rng = np.random.default_rng(1)
    u, i, j = (torch.from_numpy(rng.integers(0, n, 64)).cuda() for n in (n_users, n_items, n_items))
    mask = (torch.rand(adj.nnz) + 0.6).int().bool()
    l1 = e1.train_single_batch((u, i, j), keep_mask=mask)
    l2 = e2.train_single_batch((u, i, j), keep_mask=mask)
    assert l1 == l2
    for (k, a), (_, b) in zip(e1.model.state_dict().items(), e2.model.state_dict().items()):
        assert torch.equal(a, b), k
