"""GPU parity tests for the NCF family (GMF, MLP, NeuMF) through the C ABI:
against the reference's golden vectors and against the numpy oracle on seeded
inputs.  Tolerances as in tests/test_mf_gpu.py."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from tests.golden_util import Golden, max_rel_err, names
from tests.test_oracle_golden import BUDGET, check_adaptive_step, check_params_adaptive

pytestmark = pytest.mark.gpu


def make_engine(kind, n_users, n_items, emb_dim, n_layers, batch, optimizer, lr, state=None, adam_mode="dense"):
    from beta_recsys_b200.engines import GMFEngine, MLPEngine, NeuMFEngine

    cfg = {"model": dict(model="ncf_end", device_str="cuda:0", n_users=n_users, n_items=n_items, emb_dim=emb_dim,
                         batch_size=batch, optimizer=optimizer, lr=lr, dropout=0.0, adam_mode=adam_mode,
                         mlp_config={"n_layers": n_layers}),
           "system": {"run_dir": "/tmp/brs_test"}}
    eng = {"gmf": GMFEngine, "mlp": MLPEngine, "neumf": NeuMFEngine}[kind](cfg)
    if state is not None:
        with torch.no_grad():
            sd = eng.model.state_dict()
            assert sorted(sd) == sorted(state), (sorted(sd), sorted(state))
            for k, v in sd.items():
                v.copy_(torch.from_numpy(state[k]))
    return eng


def snap(eng):
    return {k: v.detach().cpu().numpy().copy() for k, v in eng.model.state_dict().items()}


def opt_snap(eng):
    out = {"m": {}, "v": {}}
    for name, st in eng.optimizer.state.items():
        for kind in ("m", "v"):
            if kind in st:
                out[kind][name] = st[kind].detach().cpu().numpy().copy()
    return out


def cuda(*arrs):
    return tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs)


def oracle_step(kind, p, st, u, i, r, n_layers, optimizer, lr):
    if kind == "gmf":
        return O.gmf_train_single_batch(p, st, u, i, r, optimizer=optimizer, lr=lr)
    if kind == "neumf":
        return O.neumf_train_single_batch(p, st, u, i, r, n_layers, optimizer=optimizer, lr=lr)
    return O.mlp_train_single_batch(p, st, u, i, r, n_layers, optimizer=optimizer, lr=lr)


@pytest.mark.parametrize("name", names("gmf_") + names("neumf_") + names("mlp_"))
def test_ncf_matches_reference_golden(name):
    g = Golden(name)
    m, b = g.meta, g.batch
    adaptive = m["optimizer"] in ("adam", "rmsprop")
    eng = make_engine(m["model"], m["n_users"], m["n_items"], m["emb_dim"], m["n_layers"], m["batch"], m["optimizer"],
                      m["lr"], state=g.init)
    for t in range(5):
        before, opt_before = snap(eng), opt_snap(eng)
        loss = eng.train_single_batch(*cuda(b["users"][t], b["items"][t], b["ratings"][t]))
        lt = 1e-5 if (not adaptive or t == 0) else 2e-3
        assert abs(loss - g.out["loss"][t]) <= lt * max(1, abs(g.out["loss"][t])), (t, loss, g.out["loss"][t])
        if adaptive:
            if t == 0:
                check_adaptive_step(before, snap(eng), opt_snap(eng), g.group("opt1"), m["optimizer"], m["lr"], 1)
                check_params_adaptive(snap(eng), g.group("after1"), before, m["lr"], 1)
            else:  # per-step parity from the GPU's own pre-step state
                p = {k: v.copy() for k, v in before.items()}
                st = {"step": t, "m": {k: v.copy() for k, v in opt_before["m"].items()},
                      "v": {k: v.copy() for k, v in opt_before["v"].items()}}
                ol = oracle_step(m["model"], p, st, b["users"][t], b["items"][t], b["ratings"][t], m["n_layers"],
                                 m["optimizer"], m["lr"])
                assert abs(loss - ol) <= 1e-5 * max(1, abs(ol))
                ref_opt = {f"{kind}/{k}": v for kind in ("m", "v") for k, v in st[kind].items()}
                check_adaptive_step(before, snap(eng), opt_snap(eng), ref_opt, m["optimizer"], m["lr"], t + 1)
        elif t == 0:
            for k, v in g.group("after1").items():
                assert max_rel_err(snap(eng)[k], v) <= BUDGET, (k, max_rel_err(snap(eng)[k], v))
    if not adaptive:
        for k, v in g.group("after5").items():
            assert max_rel_err(snap(eng)[k], v) <= 2 * BUDGET, (k, max_rel_err(snap(eng)[k], v))


def random_state(kind, rng, nu, ni, emb, n_layers):
    f = np.float32
    if kind == "gmf":
        return {"embedding_user.weight": rng.normal(0, 0.5, (nu, emb)).astype(f),
                "embedding_item.weight": rng.normal(0, 0.5, (ni, emb)).astype(f),
                "affine_output.weight": rng.normal(0, 0.3, (1, emb)).astype(f),
                "affine_output.bias": rng.normal(0, 0.1, (1,)).astype(f)}
    lm = emb * 2 ** (n_layers - 1)
    p = {}
    if kind == "neumf":
        p["embedding_user_mlp.weight"] = rng.normal(0, 0.5, (nu, lm)).astype(f)
        p["embedding_item_mlp.weight"] = rng.normal(0, 0.5, (ni, lm)).astype(f)
        p["embedding_user_mf.weight"] = rng.normal(0, 0.5, (nu, emb)).astype(f)
        p["embedding_item_mf.weight"] = rng.normal(0, 0.5, (ni, emb)).astype(f)
    else:
        p["embedding_user.weight"] = rng.normal(0, 0.5, (nu, lm)).astype(f)
        p["embedding_item.weight"] = rng.normal(0, 0.5, (ni, lm)).astype(f)
    for l in range(n_layers):
        fin = 2 * lm >> l
        p[f"fc_layers.{3 * l + 1}.weight"] = (rng.normal(0, 1, (fin // 2, fin)) / np.sqrt(fin)).astype(f)
        p[f"fc_layers.{3 * l + 1}.bias"] = rng.normal(0, 0.05, (fin // 2,)).astype(f)
    hw = 2 * emb if kind == "neumf" else emb
    p["affine_output.weight"] = rng.normal(0, 0.3, (1, hw)).astype(f)
    p["affine_output.bias"] = rng.normal(0, 0.1, (1,)).astype(f)
    return p


@pytest.mark.parametrize("kind,emb,n_layers", [("gmf", 64, 0), ("gmf", 32, 0), ("gmf", 8, 0), ("mlp", 16, 3),
                                               ("mlp", 32, 2), ("neumf", 16, 3), ("neumf", 64, 3), ("neumf", 32, 1),
                                               ("neumf", 8, 4)])
def test_ncf_sgd_vs_oracle(kind, emb, n_layers):
    """neumf emb=64, n_layers=3 is BASELINE config 3's model shape (MLP rows 256, tower 512->256->128->64)."""
    rng = np.random.default_rng(emb * 10 + n_layers)
    nu, ni, bsz, lr = 3000, 2000, 1500, 0.05
    p = random_state(kind, rng, nu, ni, emb, max(n_layers, 1))
    st = O.new_opt_state(p, "sgd")
    eng = make_engine(kind, nu, ni, emb, max(n_layers, 1), bsz, "sgd", lr, state=p)
    for t in range(3):
        u, i = rng.integers(0, nu, bsz), rng.integers(0, ni, bsz)
        r = (rng.random(bsz) < 0.2).astype(np.float32)
        loss = eng.train_single_batch(*cuda(u, i, r))
        ol = oracle_step(kind, p, st, u, i, r, max(n_layers, 1), "sgd", lr)
        assert abs(loss - ol) <= 1e-5 * max(1, abs(ol)), (t, loss, ol)
    got = snap(eng)
    for k in p:
        assert max_rel_err(got[k], p[k]) <= BUDGET, (k, max_rel_err(got[k], p[k]))


@pytest.mark.parametrize("kind", ["gmf", "neumf"])
def test_ncf_dense_adam_step_vs_oracle(kind):
    rng = np.random.default_rng(21)
    nu, ni, emb, nl, bsz, lr = 800, 600, 32, 2, 512, 1e-3
    p = random_state(kind, rng, nu, ni, emb, nl)
    st = O.new_opt_state(p, "adam")
    eng = make_engine(kind, nu, ni, emb, nl, bsz, "adam", lr, state=p)
    u, i = rng.integers(0, nu, bsz), rng.integers(0, ni, bsz)
    r = (rng.random(bsz) < 0.2).astype(np.float32)
    before = snap(eng)
    loss = eng.train_single_batch(*cuda(u, i, r))
    ol = oracle_step(kind, p, st, u, i, r, nl, "adam", lr)
    assert abs(loss - ol) <= 1e-5
    ref_opt = {f"{k2}/{k}": v for k2 in ("m", "v") for k, v in st[k2].items()}
    check_adaptive_step(before, snap(eng), opt_snap(eng), ref_opt, "adam", lr, 1)
    check_params_adaptive(snap(eng), p, before, lr, 1)


def test_ncf_predict_matches_oracle_forward():
    rng = np.random.default_rng(5)
    for kind, emb, nl in (("gmf", 32, 1), ("mlp", 16, 2), ("neumf", 16, 3)):
        nu, ni = 300, 200
        p = random_state(kind, rng, nu, ni, emb, nl)
        eng = make_engine(kind, nu, ni, emb, nl, 64, "sgd", 0.1, state=p)
        u, i = rng.integers(0, nu, 777), rng.integers(0, ni, 777)  # > max_batch: chunked
        s = eng.model.predict(u, i)
        assert s.shape == (777, 1)
        if kind == "gmf":
            want = O.gmf_forward(p, u, i)
        elif kind == "mlp":
            want = O.mlp_forward(p, u, i, nl)
        else:
            want = O.neumf_forward(p, u, i, nl)
        assert max_rel_err(s.cpu().numpy().ravel(), want) <= 2e-6, kind


def test_ncf_out_of_range_and_checkpoint_keys(tmp_path):
    rng = np.random.default_rng(6)
    eng = make_engine("neumf", 50, 40, 8, 2, 32, "adam", 1e-3)
    u, i = rng.integers(0, 50, 32), rng.integers(0, 40, 32)
    r = np.zeros(32, np.float32)
    bad = i.copy()
    bad[3] = 40
    with pytest.raises(IndexError):
        eng.train_single_batch(*cuda(u, bad, r))
    eng.train_single_batch(*cuda(u, i, r))
    path = str(tmp_path / "ncf.model")
    eng.save_checkpoint(path)
    keys = sorted(torch.load(path))
    assert keys == sorted(["embedding_user_mlp.weight", "embedding_item_mlp.weight", "embedding_user_mf.weight",
                           "embedding_item_mf.weight", "fc_layers.1.weight", "fc_layers.1.bias", "fc_layers.4.weight",
                           "fc_layers.4.bias", "affine_output.weight", "affine_output.bias"])


class _RatingDataset(torch.utils.data.Dataset):
    """Same shape as beta_rec.data.data_loaders.RatingDataset."""

    def __init__(self, u, i, r):
        self.user_tensor, self.item_tensor, self.target_tensor = u, i, r

    def __getitem__(self, k):
        return self.user_tensor[k], self.item_tensor[k], self.target_tensor[k]

    def __len__(self):
        return self.user_tensor.size(0)


def test_ncf_train_an_epoch_matches_loader_order():
    rng = np.random.default_rng(8)
    nu, ni, emb, nl, bsz, n = 400, 300, 16, 2, 256, 1000
    p = random_state("neumf", rng, nu, ni, emb, nl)
    u, i = rng.integers(0, nu, n), rng.integers(0, ni, n)
    r = (rng.random(n) < 0.2).astype(np.float32)
    eng = make_engine("neumf", nu, ni, emb, nl, bsz, "sgd", 0.05, state=p)
    torch.manual_seed(99)
    eng.train_an_epoch(torch.utils.data.DataLoader(_RatingDataset(*cuda(u, i, r)), batch_size=bsz, shuffle=True), 0)
    st = O.new_opt_state(p, "sgd")
    torch.manual_seed(99)
    ds = _RatingDataset(torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(r))
    for bu, bi, br in torch.utils.data.DataLoader(ds, batch_size=bsz, shuffle=True):
        O.neumf_train_single_batch(p, st, bu.numpy(), bi.numpy(), br.numpy(), nl, optimizer="sgd", lr=0.05)
    got = snap(eng)
    for k in p:
        assert max_rel_err(got[k], p[k]) <= BUDGET, (k, max_rel_err(got[k], p[k]))


@pytest.mark.parametrize("m,n,k", [(1, 8, 16), (100, 64, 128), (1000, 256, 512), (777, 33, 50), (4096, 128, 256)])
def test_linear_building_blocks_vs_numpy(m, n, k):
    """brs_mlp_fwd / brs_mlp_bwd against float64 numpy (fp32 FFMA path: ~1e-6)."""
    from beta_recsys_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(m + n + k)
    x = rng.normal(0, 1, (m, k)).astype(np.float32)
    w = (rng.normal(0, 1, (n, k)) / np.sqrt(k)).astype(np.float32)
    b = rng.normal(0, 0.1, n).astype(np.float32)
    dy = rng.normal(0, 1, (m, n)).astype(np.float32)
    tx, tw, tb, tdy = cuda(x, w, b, dy)
    ty = torch.empty((m, n), device="cuda")
    tdx = torch.empty((m, k), device="cuda")
    tdw = torch.zeros((n, k), device="cuda")
    tdb = torch.zeros(n, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.brs_mlp_fwd(tx.data_ptr(), tw.data_ptr(), tb.data_ptr(), ty.data_ptr(), m, n, k, 1, st))
    want = np.maximum(x.astype(np.float64) @ w.T.astype(np.float64) + b, 0)
    assert max_rel_err(ty.cpu().numpy(), want) <= 2e-6
    _lib.check(lib.brs_mlp_bwd(tdy.data_ptr(), tx.data_ptr(), tw.data_ptr(), tdx.data_ptr(), tdw.data_ptr(),
                               tdb.data_ptr(), tx.data_ptr(), m, n, k, st))
    want_dx = (dy.astype(np.float64) @ w.astype(np.float64)) * (x > 0)
    assert max_rel_err(tdx.cpu().numpy(), want_dx) <= 2e-6
    assert max_rel_err(tdw.cpu().numpy(), dy.T.astype(np.float64) @ x.astype(np.float64)) <= 4e-6
    assert max_rel_err(tdb.cpu().numpy(), dy.sum(0, dtype=np.float64)) <= 4e-6


@pytest.mark.parametrize("m,n,k,relu,use_mask", [(128, 64, 32, 1, 0), (1000, 256, 512, 1, 0), (4096, 128, 256, 1, 0),
                                                  (777, 64, 128, 0, 1), (65536, 256, 512, 1, 0), (300, 512, 256, 0, 1)])
def test_tensor_core_linear_3xtf32_vs_float64(m, n, k, relu, use_mask):
    """tcgen05 kind::tf32 with hi/lo split: fp32-class accuracy (plain TF32 would be ~1e-3)."""
    from beta_recsys_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(m + n + k)
    x = rng.normal(0, 1, (m, k)).astype(np.float32)
    w = (rng.normal(0, 1, (n, k)) / np.sqrt(k)).astype(np.float32)
    b = rng.normal(0, 0.1, n).astype(np.float32)
    msk = rng.normal(0, 1, (m, n)).astype(np.float32)
    tx, tw, tb, tm = cuda(x, w, b, msk)
    ty = torch.full((m, n), float("nan"), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.brs_mlp_fwd_tc(tx.data_ptr(), tw.data_ptr(), tb.data_ptr(), ty.data_ptr(),
                                  tm.data_ptr() if use_mask else None, m, n, k, relu, st), "brs_mlp_fwd_tc")
    want = x.astype(np.float64) @ w.T.astype(np.float64) + b
    if relu:
        want = np.maximum(want, 0)
    if use_mask:
        want = want * (msk > 0)
    got = ty.cpu().numpy()
    assert np.isfinite(got).all()
    assert max_rel_err(got, want) <= 5e-6, max_rel_err(got, want)


def test_neumf_same_result_on_both_gemm_backends():
    from beta_recsys_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(33)
    nu, ni, emb, nl, bsz = 2000, 1500, 64, 3, 4096
    p = random_state("neumf", rng, nu, ni, emb, nl)
    u, i = rng.integers(0, nu, bsz), rng.integers(0, ni, bsz)
    r = (rng.random(bsz) < 0.2).astype(np.float32)
    res = []
    try:
        for backend in (0, 1):
            lib.brs_set_gemm_backend(backend)
            eng = make_engine("neumf", nu, ni, emb, nl, bsz, "sgd", 0.05, state=p)
            loss = eng.train_single_batch(*cuda(u, i, r))
            res.append((loss, snap(eng)))
    finally:
        lib.brs_set_gemm_backend(1)
    assert abs(res[0][0] - res[1][0]) <= 1e-6
    for k in res[0][1]:
        assert max_rel_err(res[1][1][k], res[0][1][k]) <= BUDGET, (k, max_rel_err(res[1][1][k], res[0][1][k]))


@pytest.mark.parametrize("name", names("cfg_gmf_") + names("cfg_neumf_") + names("cfg_mlp_"))
def test_ncf_cfg_goldens_at_benchmark_dims(name):
    test_ncf_matches_reference_golden(name)
