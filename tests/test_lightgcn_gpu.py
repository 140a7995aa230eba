"""GPU parity tests for LightGCN through the C ABI: golden vectors of the reference (with
the edge-dropout masks the reference drew) and the numpy oracle on seeded graphs."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import cf_oracle as O
from tests.golden_util import Golden, max_rel_err, names
from tests.test_oracle_golden import BUDGET, check_adaptive_step, check_params_adaptive

pytestmark = pytest.mark.gpu


def make_engine(nu, ni, d, n_layers, adj_coo, optimizer, lr, decay, keep_pro, state=None, bsz=64):
    from beta_recsys_b200.engines import LightGCNEngine

    n = nu + ni
    tadj = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([adj_coo.row, adj_coo.col]).astype(np.int64)),
                                   torch.from_numpy(adj_coo.data.astype(np.float32)), (n, n))
    cfg = {"model": dict(device_str="cuda:0", n_users=nu, n_items=ni, emb_dim=d, batch_size=bsz, optimizer=optimizer,
                         lr=lr, regs=[decay], keep_pro=keep_pro, layer_size=[d] * n_layers, norm_adj=tadj),
           "system": {"run_dir": "/tmp/brs_test"}}
    eng = LightGCNEngine(cfg)
    eng.model.train()
    if state is not None:
        with torch.no_grad():
            for k, v in eng.model.state_dict().items():
                v.copy_(torch.from_numpy(state[k]))
    return eng


def snap(eng):
    return {k: v.detach().cpu().numpy().copy() for k, v in eng.model.state_dict().items()}


def cuda(*a):
    return tuple(torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in a)


@pytest.mark.parametrize("name", names("lightgcn_"))
def test_lightgcn_matches_reference_golden(name):
    g = Golden(name)
    m, b, x = g.meta, g.batch, g.group("extra")
    n = m["n_users"] + m["n_items"]
    adj = sp.coo_matrix((x["adj_val"], (x["adj_row"], x["adj_col"])), shape=(n, n))
    eng = make_engine(m["n_users"], m["n_items"], m["emb_dim"], m["n_layers"], adj, m["optimizer"], m["lr"], m["decay"],
                      m["keep_pro"], state=g.init, bsz=m["batch"])
    assert sorted(eng.model.state_dict()) == ["item_embedding.weight", "user_embedding.weight"]
    adaptive = m["optimizer"] == "adam"
    for t in range(5):
        before = snap(eng)
        loss = eng.train_single_batch(cuda(b["users"][t], b["pos"][t], b["neg"][t]), keep_mask=x["keep_masks"][t])
        lt = 1e-5 if (not adaptive or t == 0) else 2e-3
        assert abs(loss - g.out["loss"][t]) <= lt * max(1, abs(g.out["loss"][t])), (t, loss, g.out["loss"][t])
        if t == 0:
            if adaptive:
                st = eng.optimizer.state["all_embeddings"]
                mm, vv = st["m"].cpu().numpy(), st["v"].cpu().numpy()
                nu = m["n_users"]
                opt = {"m": {"user_embedding.weight": mm[:nu], "item_embedding.weight": mm[nu:]},
                       "v": {"user_embedding.weight": vv[:nu], "item_embedding.weight": vv[nu:]}}
                check_adaptive_step(before, snap(eng), opt, g.group("opt1"), "adam", m["lr"], 1)
                check_params_adaptive(snap(eng), g.group("after1"), before, m["lr"], 1)
            else:
                for k, v in g.group("after1").items():
                    assert max_rel_err(snap(eng)[k], v) <= BUDGET, (k, max_rel_err(snap(eng)[k], v))
    if not adaptive:
        for k, v in g.group("after5").items():
            assert max_rel_err(snap(eng)[k], v) <= 2 * BUDGET, (k, max_rel_err(snap(eng)[k], v))


@pytest.mark.parametrize("d,n_layers,keep", [(64, 3, 0.6), (32, 2, 1.0), (128, 1, 0.8), (16, 4, 0.5)])
def test_lightgcn_sgd_vs_oracle_skewed_graph(d, n_layers, keep):
    """Zipf-degree graph: a few rows hold thousands of non-zeros (exercises the nnz-balanced chunks)."""
    rng = np.random.default_rng(d + n_layers)
    nu, ni, n_e, bsz, lr, decay = 3000, 1500, 60000, 1024, 0.05, 1e-4
    pu = np.arange(1, nu + 1) ** -1.0
    pi = np.arange(1, ni + 1) ** -1.0
    eu = rng.choice(nu, n_e, p=pu / pu.sum())
    ei = rng.choice(ni, n_e, p=pi / pi.sum())
    adj = O.row_normalised_adj(nu, ni, eu, ei)
    p = {"user_embedding.weight": rng.normal(0, 0.1, (nu, d)).astype(np.float32),
         "item_embedding.weight": rng.normal(0, 0.1, (ni, d)).astype(np.float32)}
    st = O.new_opt_state(p, "sgd")
    eng = make_engine(nu, ni, d, n_layers, adj.tocoo(), "sgd", lr, decay, keep, state=p, bsz=bsz)
    for t in range(2):
        u, i, j = rng.integers(0, nu, bsz), rng.integers(0, ni, bsz), rng.integers(0, ni, bsz)
        mask = rng.random(adj.nnz) < keep
        loss = eng.train_single_batch(cuda(u, i, j), keep_mask=mask.astype(np.uint8))
        a = O.edge_dropout(adj, mask, keep)
        ol = O.lightgcn_train_single_batch(p, st, a, u, i, j, n_layers, decay, optimizer="sgd", lr=lr)
        assert abs(loss - ol) <= 1e-5 * max(1, abs(ol)), (t, loss, ol)
    got = snap(eng)
    for k in p:
        assert max_rel_err(got[k], p[k]) <= BUDGET, (k, max_rel_err(got[k], p[k]))


def test_spmm_building_block_and_transpose_map():
    from beta_recsys_b200 import _lib
    from beta_recsys_b200.engines.lightgcn import coo_to_csr

    lib = _lib.load()
    rng = np.random.default_rng(3)
    n, d = 700, 64
    a = sp.random(n, n, density=0.02, random_state=4, format="coo", dtype=np.float32)
    a.sum_duplicates()
    a = a.tocsr()
    a.sort_indices()
    coo = a.tocoo()
    c = coo_to_csr(coo.row, coo.col, coo.data, n)
    x = rng.normal(0, 1, (n, d)).astype(np.float32)
    mask = (rng.random(a.nnz) < 0.7).astype(np.uint8)
    t = {k: torch.from_numpy(v).cuda() for k, v in c.items() if k != "nnz"}
    tx, tm = torch.from_numpy(x).cuda(), torch.from_numpy(mask).cuda()
    st = torch.cuda.current_stream().cuda_stream
    fwd = _lib.Csr(t["row_ptr"].data_ptr(), t["col"].data_ptr(), t["val"].data_ptr(), None, n, c["nnz"])
    bwd = _lib.Csr(t["row_ptr_t"].data_ptr(), t["col_t"].data_ptr(), t["val_t"].data_ptr(), t["edge_id_t"].data_ptr(),
                   n, c["nnz"])
    y = torch.zeros((n, d), device="cuda")
    _lib.check(lib.brs_spmm_csr(fwd, tm.data_ptr(), 0.7, tx.data_ptr(), y.data_ptr(), d, st))
    am = sp.csr_matrix((coo.data * mask / np.float32(0.7), (coo.row, coo.col)), shape=(n, n))
    assert max_rel_err(y.cpu().numpy(), (am.astype(np.float64) @ x.astype(np.float64))) <= 2e-6
    yt = torch.zeros((n, d), device="cuda")
    _lib.check(lib.brs_spmm_csr(bwd, tm.data_ptr(), 0.7, tx.data_ptr(), yt.data_ptr(), d, st))
    assert max_rel_err(yt.cpu().numpy(), (am.T.astype(np.float64) @ x.astype(np.float64))) <= 2e-6


def test_lightgcn_predict_matches_oracle():
    rng = np.random.default_rng(5)
    nu, ni, d, L = 300, 200, 64, 3
    eu, ei = rng.integers(0, nu, 3000), rng.integers(0, ni, 3000)
    adj = O.row_normalised_adj(nu, ni, eu, ei)
    p = {"user_embedding.weight": rng.normal(0, 0.3, (nu, d)).astype(np.float32),
         "item_embedding.weight": rng.normal(0, 0.3, (ni, d)).astype(np.float32)}
    eng = make_engine(nu, ni, d, L, adj.tocoo(), "adam", 0.01, 1e-5, 0.6, state=p)
    u, i = rng.integers(0, nu, 500), rng.integers(0, ni, 500)
    s = eng.model.predict(u, i)
    ebar = O.lightgcn_propagate(p, adj, L)
    want = O.sigmoid((ebar[:nu][u] * ebar[nu:][i]).sum(1))
    assert max_rel_err(s.cpu().numpy(), want) <= 2e-6


@pytest.mark.parametrize("name", names("cfg_lightgcn_"))
def test_lightgcn_cfg_goldens_at_benchmark_dims(name):
    test_lightgcn_matches_reference_golden(name)
