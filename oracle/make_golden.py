"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the LIVE reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Drives the unmodified reference engines (beta_rec.models.{mf,gmf,mlp,ncf,lightgcn})
on CPU through ``oracle/ref_shim.py`` and records, for every case: the initial
``state_dict``, the index batches, and after 1 and after 5 consecutive
``train_single_batch`` calls the returned floats, every parameter tensor and
(Adam/RMSprop) the optimizer state.  These files are the parity pin for
``oracle/cf_oracle.py`` (the reference has no golden vectors of
its own for this path) and travel to the GPU box with the repo.
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
N_STEPS = 5


def _quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _snap(model):
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def _opt_state(engine):
    """exp_avg / exp_avg_sq / square_avg per parameter name."""
    import torch  # noqa: F401

    names = [n for n, _ in engine.model.named_parameters()]
    params = [p for _, p in engine.model.named_parameters()]
    out = {}
    for n, p in zip(names, params):
        st = engine.optimizer.state.get(p, {})
        for src, dst in (("exp_avg", "m"), ("exp_avg_sq", "v"), ("square_avg", "v")):
            if src in st:
                out[f"{dst}/{n}"] = st[src].detach().cpu().numpy().copy()
    return out


def _save(name, meta, init, batches, floats, after1, after5, opt5, extra=None, opt1=None):
    d = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
    for k, v in init.items():
        d["init/" + k] = v
    for k, v in batches.items():
        d["batch/" + k] = v
    for k, v in floats.items():
        d["out/" + k] = np.asarray(v, dtype=np.float64)
    for k, v in after1.items():
        d["after1/" + k] = v
    for k, v in after5.items():
        d["after5/" + k] = v
    for k, v in opt5.items():
        d["opt5/" + k] = v
    for k, v in (opt1 or {}).items():
        d["opt1/" + k] = v
    for k, v in (extra or {}).items():
        d["extra/" + k] = v
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name, "%.0f KB" % (os.path.getsize(os.path.join(OUT, name + ".npz")) / 1024))


def _idx_batches(rng, n_users, n_items, b, mode):
    if mode == "random":
        u = rng.integers(0, n_users, (N_STEPS, b))
        p = rng.integers(0, n_items, (N_STEPS, b))
        n = rng.integers(0, n_items, (N_STEPS, b))
    elif mode == "dup":  # every triple shares one user; items drawn from 3 ids
        u = np.repeat(rng.integers(0, n_users, (N_STEPS, 1)), b, axis=1)
        p = rng.integers(0, 3, (N_STEPS, b))
        n = rng.integers(0, 3, (N_STEPS, b))
    elif mode == "nodup":  # no row touched twice inside a batch
        assert 2 * b <= n_items and b <= n_users
        u = np.stack([rng.permutation(n_users)[:b] for _ in range(N_STEPS)])
        pi = np.stack([rng.permutation(n_items)[: 2 * b] for _ in range(N_STEPS)])
        p, n = pi[:, :b], pi[:, b:]
    else:
        raise ValueError(mode)
    return u.astype(np.int64), p.astype(np.int64), n.astype(np.int64)


def gen_mf(name, d, optimizer, loss, mode="random", n_users=48, n_items=40, b=32, lr=0.05, seed=0, f64=False):
    import torch
    from beta_rec.models.mf import MFEngine

    torch.manual_seed(seed)
    cfg = {
        "model": dict(device_str="cpu", n_users=n_users, n_items=n_items, emb_dim=d, batch_size=b,
                      optimizer=optimizer, lr=lr, loss=loss, reg=0.001),
        "system": {"run_dir": "/tmp/brs_golden"},
    }
    eng = _quiet(MFEngine, cfg)
    assert eng.reg == 0.0  # beta_rec/models/mf.py:81-83 looks at the top-level config
    with torch.no_grad():  # non-trivial biases so their gradients are exercised
        eng.model.user_bias.weight.normal_(0, 0.1)
        eng.model.item_bias.weight.normal_(0, 0.1)
        eng.model.global_bias.fill_(0.05)
    if f64:  # same Parameter objects -> the optimizer built in __init__ keeps tracking them
        eng.model.double()
    init = _snap(eng.model)
    rng = np.random.default_rng(seed + 1)
    u, p, n = _idx_batches(rng, n_users, n_items, b, mode)
    r = (rng.random((N_STEPS, b)) < 0.4).astype(np.float64 if f64 else np.float32)
    losses, regs = [], []
    after1 = opt1 = None
    for t in range(N_STEPS):
        if loss == "bpr":
            batch = (torch.from_numpy(u[t]), torch.from_numpy(p[t]), torch.from_numpy(n[t]))
        else:
            batch = (torch.from_numpy(u[t]), torch.from_numpy(p[t]), torch.from_numpy(r[t]))
        l, g = eng.train_single_batch(batch)
        losses.append(l)
        regs.append(g)
        if t == 0:
            after1, opt1 = _snap(eng.model), _opt_state(eng)
    meta = dict(model="mf", emb_dim=d, optimizer=optimizer, loss=loss, lr=lr, n_users=n_users, n_items=n_items,
                batch=b, mode=mode, reg=0.0, torch=torch.__version__, dtype="f64" if f64 else "f32")
    _save(name, meta, init, {"users": u, "pos": p, "neg": n, "ratings": r}, {"loss": losses, "reg": regs},
          after1, _snap(eng.model), _opt_state(eng), opt1=opt1)


def gen_ncf(name, kind, emb_dim, n_layers, optimizer, n_users=40, n_items=36, b=32, lr=1e-3, seed=0, f64=False):
    import torch
    from beta_rec.models.gmf import GMFEngine
    from beta_rec.models.mlp import MLPEngine
    from beta_rec.models.ncf import NeuMFEngine

    torch.manual_seed(seed)
    cfg = {
        "model": dict(model="ncf_end", device_str="cpu", n_users=n_users, n_items=n_items, emb_dim=emb_dim,
                      batch_size=b, optimizer=optimizer, lr=lr, dropout=0.0,
                      mlp_config={"n_layers": n_layers}),
        "system": {"run_dir": "/tmp/brs_golden"},
    }
    eng = _quiet({"gmf": GMFEngine, "mlp": MLPEngine, "neumf": NeuMFEngine}[kind], cfg)
    with torch.no_grad():  # larger-than-default embeddings so ReLU masks are non-trivial
        for n_, p_ in eng.model.named_parameters():
            if "embedding" in n_:
                p_.normal_(0, 0.5)
    if f64:
        eng.model.double()
    init = _snap(eng.model)
    rng = np.random.default_rng(seed + 1)
    u = rng.integers(0, n_users, (N_STEPS, b)).astype(np.int64)
    i = rng.integers(0, n_items, (N_STEPS, b)).astype(np.int64)
    r = (rng.random((N_STEPS, b)) < 0.3).astype(np.float64 if f64 else np.float32)
    losses, after1, opt1 = [], None, None
    for t in range(N_STEPS):
        losses.append(eng.train_single_batch(torch.from_numpy(u[t]), torch.from_numpy(i[t]), torch.from_numpy(r[t])))
        if t == 0:
            after1, opt1 = _snap(eng.model), _opt_state(eng)
    meta = dict(model=kind, emb_dim=emb_dim, n_layers=n_layers, optimizer=optimizer, lr=lr, n_users=n_users,
                n_items=n_items, batch=b, torch=torch.__version__, dtype="f64" if f64 else "f32")
    _save(name, meta, init, {"users": u, "items": i, "ratings": r}, {"loss": losses}, after1, _snap(eng.model),
          _opt_state(eng), opt1=opt1)


def gen_lightgcn(name, d, n_layers, optimizer, keep_pro=0.6, n_users=40, n_items=30, n_edges=260, b=32,
                 lr=0.05, decay=1e-5, seed=0, f64=False):
    import scipy.sparse as sp
    import torch
    from beta_rec.models.lightgcn import LightGCNEngine
    from beta_rec.utils.common_util import normalized_adj_single

    torch.manual_seed(seed)
    rng = np.random.default_rng(seed + 1)
    pairs = rng.permutation(n_users * n_items)[:n_edges]
    eu, ei = (pairs // n_items).astype(np.int64), (pairs % n_items).astype(np.int64)
    # create_adj_mat (beta_rec/data/base_data.py:337-360) without the BaseData wrapper
    n = n_users + n_items
    rmat = sp.csr_matrix((np.ones(n_edges, dtype=np.float32), (eu, ei)), shape=(n_users, n_items))
    adj = sp.bmat([[None, rmat], [rmat.T, None]], format="csr", dtype=np.float32)
    norm = _quiet(normalized_adj_single, adj + sp.eye(n)).tocsr().tocoo().astype(np.float32)
    # sparse_mx_to_torch_sparse_tensor (beta_rec/recommenders/lightgcn.py:15-23)
    idx = torch.from_numpy(np.vstack((norm.row, norm.col)).astype(np.int64))
    vals = torch.from_numpy(norm.data)
    tadj = torch.sparse_coo_tensor(idx, vals.double() if f64 else vals, torch.Size(norm.shape))
    cfg = {
        "model": dict(device_str="cpu", n_users=n_users, n_items=n_items, emb_dim=d, batch_size=b,
                      optimizer=optimizer, lr=lr, regs=[decay], keep_pro=keep_pro, layer_size=[d] * n_layers,
                      norm_adj=tadj),
        "system": {"run_dir": "/tmp/brs_golden"},
    }
    eng = _quiet(LightGCNEngine, cfg)
    eng.model.train()
    if f64:
        eng.model.double()
    init = _snap(eng.model)
    coal = tadj.coalesce()
    nnz = coal.values().numel()
    u = rng.integers(0, n_users, (N_STEPS, b)).astype(np.int64)
    p = rng.integers(0, n_items, (N_STEPS, b)).astype(np.int64)
    ng = rng.integers(0, n_items, (N_STEPS, b)).astype(np.int64)
    masks = np.zeros((N_STEPS, nnz), dtype=np.uint8)
    losses, after1, opt1 = [], None, None
    for t in range(N_STEPS):
        # LightGCN.dropout draws torch.rand(nnz) from the global CPU generator first thing in forward
        torch.manual_seed(1000 + t)
        masks[t] = (torch.rand(nnz) + keep_pro).int().bool().numpy()
        torch.manual_seed(1000 + t)
        losses.append(eng.train_single_batch((torch.from_numpy(u[t]), torch.from_numpy(p[t]), torch.from_numpy(ng[t]))))
        if t == 0:
            after1, opt1 = _snap(eng.model), _opt_state(eng)
    meta = dict(model="lightgcn", emb_dim=d, n_layers=n_layers, optimizer=optimizer, lr=lr, keep_pro=keep_pro,
                decay=decay, n_users=n_users, n_items=n_items, batch=b, torch=torch.__version__,
                dtype="f64" if f64 else "f32")
    extra = {
        "edge_users": eu, "edge_items": ei,
        "adj_row": coal.indices()[0].numpy(), "adj_col": coal.indices()[1].numpy(), "adj_val": coal.values().numpy(),
        "keep_masks": masks,
    }
    _save(name, meta, init, {"users": u, "pos": p, "neg": ng}, {"loss": losses}, after1, _snap(eng.model),
          _opt_state(eng), extra, opt1=opt1)


def gen_bench_shapes():
    """The shapes BASELINE.json's configs 2-4 run at (D = 128 MF, emb 64 / 3-layer NeuMF tower
    512-256-128-64, D = 128 LightGCN), on small tables.  Prefix cfg_: pinned by the CPU oracle tests."""
    gen_mf("cfg_mf_bce_d128_adam", 128, "adam", "bce")
    gen_mf("cfg_mf_bpr_d128_rmsprop", 128, "rmsprop", "bpr", lr=0.01)
    gen_ncf("cfg_gmf_d128_sgd", "gmf", 128, 3, "sgd", lr=0.01)
    for opt in ("sgd", "adam"):
        gen_ncf(f"cfg_neumf_e64_l3_{opt}", "neumf", 64, 3, opt, lr=0.01)
    gen_ncf("cfg_mlp_e64_l3_sgd", "mlp", 64, 3, "sgd", lr=0.01)
    gen_lightgcn("cfg_lightgcn_d128_l3_adam", 128, 3, "adam")


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_shim

    ref_shim.install()
    if "--bench-shapes-only" in sys.argv:  # add the cfg_* files without regenerating the rest
        return gen_bench_shapes()
    gen_bench_shapes()
    for d in (64, 128):
        for opt in ("sgd", "adam"):
            gen_mf(f"mf_bpr_d{d}_{opt}", d, opt, "bpr")
    gen_mf("mf_bpr_d64_rmsprop", 64, "rmsprop", "bpr", lr=0.01)
    gen_mf("mf_bpr_d32_adam_dup", 32, "adam", "bpr", mode="dup")
    gen_mf("mf_bpr_d128_sgd_dup", 128, "sgd", "bpr", mode="dup")
    gen_mf("mf_bpr_d64_sgd_nodup", 64, "sgd", "bpr", mode="nodup", n_users=48, n_items=40, b=16)
    gen_mf("mf_bpr_d256_sgd", 256, "sgd", "bpr", n_users=24, n_items=20, b=16)
    for opt in ("sgd", "adam"):
        gen_mf(f"mf_bce_d64_{opt}", 64, opt, "bce")
    for opt in ("sgd", "adam"):
        gen_ncf(f"gmf_d64_{opt}", "gmf", 64, 3, opt, lr=0.01)
        gen_ncf(f"neumf_e16_l3_{opt}", "neumf", 16, 3, opt, lr=0.01)
    gen_ncf("neumf_e32_l2_adam", "neumf", 32, 2, "adam", n_users=24, n_items=20, b=16, lr=0.01)
    gen_ncf("mlp_e16_l3_adam", "mlp", 16, 3, "adam", lr=0.01)
    for opt in ("sgd", "adam"):
        gen_lightgcn(f"lightgcn_d64_l3_{opt}", 64, 3, opt)
    gen_lightgcn("lightgcn_d32_l2_adam_nodrop", 32, 2, "adam", keep_pro=1.0)
    # float64 runs of the SAME reference code: pin the math of the adaptive
    # optimizers where fp32 comparisons are ill-conditioned (g ~ eps)
    gen_mf("f64_mf_bpr_d64_adam", 64, "adam", "bpr", f64=True)
    gen_mf("f64_mf_bpr_d32_rmsprop", 32, "rmsprop", "bpr", lr=0.01, f64=True)
    gen_mf("f64_mf_bce_d32_adam", 32, "adam", "bce", f64=True)
    gen_ncf("f64_gmf_d32_adam", "gmf", 32, 3, "adam", lr=0.01, f64=True)
    gen_ncf("f64_neumf_e16_l2_adam", "neumf", 16, 2, "adam", lr=0.01, f64=True)
    gen_ncf("f64_mlp_e16_l2_adam", "mlp", 16, 2, "adam", lr=0.01, f64=True)
    gen_lightgcn("f64_lightgcn_d32_l3_adam", 32, 3, "adam", f64=True)


if __name__ == "__main__":
    main()
