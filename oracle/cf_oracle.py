"""TEST INFRASTRUCTURE ONLY -- numpy restatement of beta_rec's embedding-CF training step.

Closed-form forward, hand-derived gradients and explicit optimizer steps (no
autograd, no torch) for the models BASELINE.json names.  Every function cites
the reference file:line it follows (paths relative to /root/reference).  The
arithmetic the reference delegates to PyTorch (third-party, pinned only as
``torch>=1.7.1`` in requirements.txt:4; 2.11.0 in this image) is restated from
its published formulas: ``torch/optim/adam.py`` ``_single_tensor_adam``,
``torch/optim/sgd.py``, ``torch/optim/rmsprop.py``, ``F.logsigmoid``,
``nn.BCELoss`` (log clamped at -100), ``F.softplus`` (threshold 20).

Pinned by ``tests/test_oracle_golden.py`` against outputs of the reference
itself (``tests/golden/*.npz``, produced by ``oracle/make_golden.py`` through
``oracle/ref_shim.py``).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module.

State layout: a dict of numpy arrays keyed exactly like the reference
module's ``state_dict()``; optimizer state is ``{"step": int, "m": {...},
"v": {...}}`` keyed the same way.
"""
import numpy as np
import scipy.sparse as sp

F32 = np.float32


# --------------------------------------------------------------------------- #
# scalar helpers (formulas PyTorch applies at the reference call sites)
# --------------------------------------------------------------------------- #
def sigmoid(x):
    """torch.sigmoid (beta_rec/models/mf.py:43, gmf.py:35, ncf.py:70)."""
    x = np.asarray(x)
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1 / (1 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1 + e)
    return out


def logsigmoid(x):
    """F.logsigmoid = min(x,0) - log1p(exp(-|x|)) (beta_rec/models/torch_engine.py:104)."""
    return np.minimum(x, 0) - np.log1p(np.exp(-np.abs(x)))


def softplus(x):
    """F.softplus, beta=1, threshold=20 (beta_rec/models/lightgcn.py:189)."""
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))


def _scatter_rows(idx, rows, n):
    """sum rows[k] into out[idx[k]] -- what embedding_dense_backward does for
    nn.Embedding(sparse=False) (beta_rec/models/mf.py:21-24)."""
    idx = np.asarray(idx, dtype=np.int64)
    rows = np.asarray(rows)
    b = idx.shape[0]
    sel = sp.csr_matrix(
        (np.ones(b, dtype=rows.dtype), (idx, np.arange(b, dtype=np.int64))), shape=(n, b)
    )
    out = sel @ rows.reshape(b, -1)
    return np.asarray(out).reshape((n,) + rows.shape[1:])


# --------------------------------------------------------------------------- #
# optimizers (beta_rec/models/torch_engine.py:23-39 -> torch.optim defaults)
# --------------------------------------------------------------------------- #
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8
RMS_ALPHA, RMS_EPS = 0.99, 1e-8


def new_opt_state(params, optimizer):
    st = {"step": 0}
    if optimizer == "adam":
        st["m"] = {k: np.zeros_like(v) for k, v in params.items()}
        st["v"] = {k: np.zeros_like(v) for k, v in params.items()}
    elif optimizer == "rmsprop":
        st["v"] = {k: np.zeros_like(v) for k, v in params.items()}
    return st


def optimizer_step(params, grads, st, optimizer, lr):
    """One dense torch.optim step over ALL parameters (rows with zero gradient
    included: Adam keeps moving them through the decaying first moment --
    SURVEY.md section 0.6).  Updates ``params``/``st`` in place."""
    st["step"] += 1
    t = st["step"]
    for k, p in params.items():
        g = grads[k].astype(p.dtype)
        dt = p.dtype.type
        if optimizer == "sgd":  # torch/optim/sgd.py: p.add_(g, alpha=-lr)
            p -= dt(lr) * g
        elif optimizer == "adam":  # torch/optim/adam.py _single_tensor_adam
            m, v = st["m"][k], st["v"][k]
            m += (g - m) * dt(1 - ADAM_B1)  # exp_avg.lerp_(grad, 1-beta1)
            v *= dt(ADAM_B2)
            v += dt(1 - ADAM_B2) * g * g  # mul_(beta2).addcmul_(g,g,1-beta2)
            bc1 = 1 - ADAM_B1 ** t
            bc2 = 1 - ADAM_B2 ** t
            step_size = lr / bc1
            denom = np.sqrt(v) / dt(bc2 ** 0.5) + dt(ADAM_EPS)
            p -= dt(step_size) * (m / denom)  # addcdiv_(m, denom, -step_size)
        elif optimizer == "rmsprop":  # torch/optim/rmsprop.py, momentum=0, centered=False
            v = st["v"][k]
            v *= dt(RMS_ALPHA)
            v += dt(1 - RMS_ALPHA) * g * g
            p -= dt(lr) * (g / (np.sqrt(v) + dt(RMS_EPS)))
        else:
            raise ValueError(optimizer)


# --------------------------------------------------------------------------- #
# MF  (beta_rec/models/mf.py)
# --------------------------------------------------------------------------- #
MF_KEYS = ("global_bias", "user_emb.weight", "item_emb.weight", "user_bias.weight", "item_bias.weight")


def mf_forward(p, users, items):
    """MF.forward (beta_rec/models/mf.py:32-55): sigmoid score and the
    per-call regularizer (sum u^2 + sum i^2 + sum bu^2 + sum bi^2)/B."""
    u = p["user_emb.weight"][users]
    i = p["item_emb.weight"][items]
    bu = p["user_bias.weight"][users, 0]
    bi = p["item_bias.weight"][items, 0]
    z = (u * i).sum(1) + bu + bi + p["global_bias"][0]
    s = sigmoid(z)
    reg = ((u ** 2).sum() + (i ** 2).sum() + (bu ** 2).sum() + (bi ** 2).sum()) / u.dtype.type(u.shape[0])
    return s, reg


def _mf_zero_grads(p):
    return {k: np.zeros_like(v) for k, v in p.items()}


def _mf_accumulate(p, g, users, items, cz, reg_w, b):
    """Add d(loss)/d(params) for one MF.forward call given cz = dL/dz per sample,
    plus reg_w * d(regularizer)/d(params) (beta_rec/models/mf.py:49-54,116)."""
    dt = p["user_emb.weight"].dtype.type
    nu, ni = p["user_emb.weight"].shape[0], p["item_emb.weight"].shape[0]
    u = p["user_emb.weight"][users]
    i = p["item_emb.weight"][items]
    bu = p["user_bias.weight"][users, 0]
    bi = p["item_bias.weight"][items, 0]
    two_over_b = dt(2.0 * reg_w / b)
    g["user_emb.weight"] += _scatter_rows(users, cz[:, None] * i + two_over_b * u, nu)
    g["item_emb.weight"] += _scatter_rows(items, cz[:, None] * u + two_over_b * i, ni)
    g["user_bias.weight"][:, 0] += _scatter_rows(users, cz + two_over_b * bu, nu)
    g["item_bias.weight"][:, 0] += _scatter_rows(items, cz + two_over_b * bi, ni)
    g["global_bias"][0] += cz.sum()


def bce_loss_and_grad(s, r):
    """nn.BCELoss(reduction='mean') value and d(loss)/d(s)
    (beta_rec/models/torch_engine.py:119-120).  torch clamps each log at -100
    and computes the backward as (s-r)/max((1-s)*s, 1e-12)/B."""
    dt = s.dtype.type
    b = s.shape[0]
    log_s = np.maximum(np.log(s), dt(-100))
    log_1ms = np.maximum(np.log1p(-s), dt(-100))
    loss = (-(r * log_s + (1 - r) * log_1ms)).mean()
    ds = (s - r) / np.maximum((1 - s) * s, dt(1e-12)) / dt(b)
    return loss, ds


def mf_bpr_loss_grads(p, users, pos, neg, reg_w=0.0):
    """MFEngine.train_single_batch, loss == 'bpr' (beta_rec/models/mf.py:102-107,116-117)
    with ModelEngine.bpr_loss (beta_rec/models/torch_engine.py:104-105).  NOTE the
    reference applies sigmoid to each score BEFORE the BPR difference."""
    dt = p["user_emb.weight"].dtype.type
    b = len(users)
    sp_, reg_p = mf_forward(p, users, pos)
    sn_, reg_n = mf_forward(p, users, neg)
    x = sp_ - sn_
    loss = -logsigmoid(x).mean()
    dx = -sigmoid(-x) / dt(b)  # d/dx of -mean(logsigmoid(x))
    cz_p = dx * sp_ * (1 - sp_)
    cz_n = -dx * sn_ * (1 - sn_)
    g = _mf_zero_grads(p)
    _mf_accumulate(p, g, users, pos, cz_p, reg_w, b)
    _mf_accumulate(p, g, users, neg, cz_n, reg_w, b)
    return loss, reg_p + reg_n, g


def mf_bce_loss_grads(p, users, items, ratings, reg_w=0.0):
    """MFEngine.train_single_batch, loss == 'bce' (beta_rec/models/mf.py:108-111)."""
    b = len(users)
    s, reg = mf_forward(p, users, items)
    loss, ds = bce_loss_and_grad(s, ratings.astype(s.dtype))
    cz = ds * s * (1 - s)
    g = _mf_zero_grads(p)
    _mf_accumulate(p, g, users, items, cz, reg_w, b)
    return loss, reg, g


def mf_condition_scale(p, batch, loss="bpr"):
    """Per parameter tensor, max over elements of sum_k |term_k| of its gradient sum
    (the forward-error scale of an fp32 summation: |fl(sum) - sum| <= c*eps*sum|terms|).
    Bias gradients add opposite-sign terms (c_pos < 0 < c_neg) and can cancel to ~0,
    so parity of such entries is judged against this scale, not against |result|."""
    pa = {k: np.abs(v) for k, v in p.items()}
    g = _mf_zero_grads(p)
    b = len(batch[0])
    dt = p["user_emb.weight"].dtype.type
    if loss == "bpr":
        users, pos, neg = batch
        sp_, _ = mf_forward(p, users, pos)
        sn_, _ = mf_forward(p, users, neg)
        dx = sigmoid(-(sp_ - sn_)) / dt(b)
        _mf_accumulate(pa, g, users, pos, np.abs(dx * sp_ * (1 - sp_)), 0.0, b)
        _mf_accumulate(pa, g, users, neg, np.abs(dx * sn_ * (1 - sn_)), 0.0, b)
    else:
        users, items, ratings = batch
        s, _ = mf_forward(p, users, items)
        _, ds = bce_loss_and_grad(s, ratings.astype(s.dtype))
        _mf_accumulate(pa, g, users, items, np.abs(ds * s * (1 - s)), 0.0, b)
    return {k: float(np.abs(v).max()) for k, v in g.items()}


def mf_train_single_batch(p, st, batch, loss="bpr", optimizer="sgd", lr=0.05, reg_w=0.0):
    """One MFEngine.train_single_batch (beta_rec/models/mf.py:92-119): returns
    (loss, regularizer) and updates p/st in place, batch-synchronously."""
    if loss == "bpr":
        l, r, g = mf_bpr_loss_grads(p, *batch, reg_w=reg_w)
    elif loss == "bce":
        l, r, g = mf_bce_loss_grads(p, *batch, reg_w=reg_w)
    else:
        raise RuntimeError(f"Unsupported loss type {loss}, try other options: 'bpr' or 'bce'")
    optimizer_step(p, g, st, optimizer, lr)
    return float(l), float(r)


# --------------------------------------------------------------------------- #
# GMF  (beta_rec/models/gmf.py)
# --------------------------------------------------------------------------- #
def gmf_forward(p, users, items):
    """GMF.forward (beta_rec/models/gmf.py:29-36): sigmoid(w . (u*i) + b)."""
    u = p["embedding_user.weight"][users]
    i = p["embedding_item.weight"][items]
    z = (u * i) @ p["affine_output.weight"][0] + p["affine_output.bias"][0]
    return sigmoid(z)


def gmf_loss_grads(p, users, items, ratings):
    """GMFEngine.train_single_batch (beta_rec/models/gmf.py:60-80), BCELoss."""
    nu, ni = p["embedding_user.weight"].shape[0], p["embedding_item.weight"].shape[0]
    u = p["embedding_user.weight"][users]
    i = p["embedding_item.weight"][items]
    w = p["affine_output.weight"][0]
    s = gmf_forward(p, users, items)
    loss, ds = bce_loss_and_grad(s, ratings.astype(s.dtype))
    dz = ds * s * (1 - s)
    g = {
        "embedding_user.weight": _scatter_rows(users, dz[:, None] * (w * i), nu),
        "embedding_item.weight": _scatter_rows(items, dz[:, None] * (w * u), ni),
        "affine_output.weight": (dz[:, None] * (u * i)).sum(0)[None, :],
        "affine_output.bias": np.array([dz.sum()], dtype=s.dtype),
    }
    return loss, g


def gmf_train_single_batch(p, st, users, items, ratings, optimizer="adam", lr=1e-3):
    l, g = gmf_loss_grads(p, users, items, ratings)
    optimizer_step(p, g, st, optimizer, lr)
    return float(l)


# --------------------------------------------------------------------------- #
# MLP tower shared by MLP and NeuMF
# --------------------------------------------------------------------------- #
def _tower_forward(x, weights, biases, neumf_quirk):
    """fc_layers loop.  neumf_quirk=True reproduces NeuMF.forward's extra ReLU
    after EVERY sub-module including the leading Dropout (beta_rec/models/ncf.py:64-66),
    i.e. the concatenated embeddings are ReLU'd before the first Linear.
    MLP.forward (beta_rec/models/mlp.py:47-48) has no such extra ReLU.
    Dropout p=0 / eval is the identity.  Returns activations list (inputs to each Linear) and output."""
    acts = []
    if neumf_quirk:
        x = np.maximum(x, 0)
    for w, b in zip(weights, biases):
        acts.append(x)
        x = np.maximum(x @ w.T + b, 0)
    return acts, x


def _tower_backward(dout, out, acts, weights, neumf_quirk, x_in):
    """Backward through the tower; returns (dx_in, [dW], [db])."""
    dws, dbs = [None] * len(weights), [None] * len(weights)
    d = dout
    y = out
    for l in range(len(weights) - 1, -1, -1):
        d = d * (y > 0)  # ReLU (applied once or twice: same mask)
        dws[l] = d.T @ acts[l]
        dbs[l] = d.sum(0)
        d = d @ weights[l]
        y = acts[l]
    if neumf_quirk:
        d = d * (x_in > 0)
    return d, dws, dbs


def _fc_keys(n_layers):
    # nn.Sequential of [Dropout, Linear, ReLU] * n -> Linear modules at 1, 4, 7, ...
    return [f"fc_layers.{3 * l + 1}" for l in range(n_layers)]


# --------------------------------------------------------------------------- #
# NeuMF  (beta_rec/models/ncf.py)
# --------------------------------------------------------------------------- #
def neumf_forward(p, users, items, n_layers, want_cache=False):
    """NeuMF.forward (beta_rec/models/ncf.py:52-71)."""
    um = p["embedding_user_mlp.weight"][users]
    im = p["embedding_item_mlp.weight"][items]
    uf = p["embedding_user_mf.weight"][users]
    if_ = p["embedding_item_mf.weight"][items]
    x0 = np.concatenate([um, im], axis=1)
    ws = [p[k + ".weight"] for k in _fc_keys(n_layers)]
    bs = [p[k + ".bias"] for k in _fc_keys(n_layers)]
    acts, h = _tower_forward(x0, ws, bs, neumf_quirk=True)
    vec = np.concatenate([h, uf * if_], axis=1)
    z = vec @ p["affine_output.weight"][0] + p["affine_output.bias"][0]
    s = sigmoid(z)
    if want_cache:
        return s, (um, im, uf, if_, x0, ws, acts, h, vec)
    return s


def neumf_loss_grads(p, users, items, ratings, n_layers):
    """NeuMFEngine.train_single_batch (beta_rec/models/ncf.py:100-120), BCELoss."""
    nu, ni = p["embedding_user_mlp.weight"].shape[0], p["embedding_item_mlp.weight"].shape[0]
    s, (um, im, uf, if_, x0, ws, acts, h, vec) = neumf_forward(p, users, items, n_layers, True)
    loss, ds = bce_loss_and_grad(s, ratings.astype(s.dtype))
    dz = ds * s * (1 - s)
    wo = p["affine_output.weight"][0]
    dvec = dz[:, None] * wo[None, :]
    hd = h.shape[1]
    dx0, dws, dbs = _tower_backward(dvec[:, :hd], h, acts, ws, True, x0)
    dmf = dvec[:, hd:]
    lm = um.shape[1]
    g = {
        "embedding_user_mlp.weight": _scatter_rows(users, dx0[:, :lm], nu),
        "embedding_item_mlp.weight": _scatter_rows(items, dx0[:, lm:], ni),
        "embedding_user_mf.weight": _scatter_rows(users, dmf * if_, nu),
        "embedding_item_mf.weight": _scatter_rows(items, dmf * uf, ni),
        "affine_output.weight": (dz[:, None] * vec).sum(0)[None, :],
        "affine_output.bias": np.array([dz.sum()], dtype=s.dtype),
    }
    for k, dw, db in zip(_fc_keys(n_layers), dws, dbs):
        g[k + ".weight"] = dw
        g[k + ".bias"] = db
    return loss, g


def neumf_train_single_batch(p, st, users, items, ratings, n_layers, optimizer="adam", lr=1e-3):
    l, g = neumf_loss_grads(p, users, items, ratings, n_layers)
    optimizer_step(p, g, st, optimizer, lr)
    return float(l)


# --------------------------------------------------------------------------- #
# MLP  (beta_rec/models/mlp.py)
# --------------------------------------------------------------------------- #
def mlp_forward(p, users, items, n_layers, want_cache=False):
    """MLP.forward (beta_rec/models/mlp.py:40-51)."""
    um = p["embedding_user.weight"][users]
    im = p["embedding_item.weight"][items]
    x0 = np.concatenate([um, im], axis=1)
    ws = [p[k + ".weight"] for k in _fc_keys(n_layers)]
    bs = [p[k + ".bias"] for k in _fc_keys(n_layers)]
    acts, h = _tower_forward(x0, ws, bs, neumf_quirk=False)
    z = h @ p["affine_output.weight"][0] + p["affine_output.bias"][0]
    s = sigmoid(z)
    if want_cache:
        return s, (um, im, x0, ws, acts, h)
    return s


def mlp_loss_grads(p, users, items, ratings, n_layers):
    """MLPEngine.train_single_batch (beta_rec/models/mlp.py:75-98)."""
    nu, ni = p["embedding_user.weight"].shape[0], p["embedding_item.weight"].shape[0]
    s, (um, im, x0, ws, acts, h) = mlp_forward(p, users, items, n_layers, True)
    loss, ds = bce_loss_and_grad(s, ratings.astype(s.dtype))
    dz = ds * s * (1 - s)
    wo = p["affine_output.weight"][0]
    dx0, dws, dbs = _tower_backward(dz[:, None] * wo[None, :], h, acts, ws, False, x0)
    lm = um.shape[1]
    g = {
        "embedding_user.weight": _scatter_rows(users, dx0[:, :lm], nu),
        "embedding_item.weight": _scatter_rows(items, dx0[:, lm:], ni),
        "affine_output.weight": (dz[:, None] * h).sum(0)[None, :],
        "affine_output.bias": np.array([dz.sum()], dtype=s.dtype),
    }
    for k, dw, db in zip(_fc_keys(n_layers), dws, dbs):
        g[k + ".weight"] = dw
        g[k + ".bias"] = db
    return loss, g


def mlp_train_single_batch(p, st, users, items, ratings, n_layers, optimizer="adam", lr=1e-3):
    l, g = mlp_loss_grads(p, users, items, ratings, n_layers)
    optimizer_step(p, g, st, optimizer, lr)
    return float(l)


# --------------------------------------------------------------------------- #
# LightGCN  (beta_rec/models/lightgcn.py)
# --------------------------------------------------------------------------- #
def row_normalised_adj(n_users, n_items, users, items, dtype=F32):
    """Ahat = D^-1 (A + I) in CSR: create_adj_mat + normalized_adj_single
    (beta_rec/data/base_data.py:337-360, beta_rec/utils/common_util.py:24-41).
    Row-normalised, hence asymmetric -- not the paper's D^-1/2 A D^-1/2."""
    n = n_users + n_items
    r = sp.csr_matrix((np.ones(len(users), dtype=dtype), (users, items)), shape=(n_users, n_items))
    r.data[:] = 1
    a = sp.bmat([[None, r], [r.T, None]], format="csr", dtype=dtype) + sp.eye(n, dtype=dtype, format="csr")
    rowsum = np.asarray(a.sum(1)).ravel()
    with np.errstate(divide="ignore"):
        d_inv = np.power(rowsum, -1)
    d_inv[np.isinf(d_inv)] = 0.0
    out = (sp.diags(d_inv.astype(dtype)) @ a).tocsr()
    out.sort_indices()
    return out.astype(dtype)


def edge_dropout(adj_csr, keep_mask, keep_prob):
    """LightGCN.dropout (beta_rec/models/lightgcn.py:27-38) given the boolean
    keep mask the reference draws as ``(torch.rand(nnz)+keep_prob).int().bool()``
    over the COALESCED (row-major sorted) edge list; kept values are / keep_prob."""
    coo = adj_csr.tocoo()  # CSR with sorted indices -> row-major order == coalesce order
    keep = np.asarray(keep_mask, dtype=bool)
    vals = (coo.data[keep] / coo.data.dtype.type(keep_prob)).astype(coo.data.dtype)
    out = sp.csr_matrix((vals, (coo.row[keep], coo.col[keep])), shape=adj_csr.shape)
    out.sort_indices()
    return out


def lightgcn_propagate(p, adj, n_layers):
    """LightGCN.forward (beta_rec/models/lightgcn.py:46-78): E^(l+1) = Ahat E^(l),
    output = mean over the L+1 layer embeddings."""
    e = np.concatenate([p["user_embedding.weight"], p["item_embedding.weight"]], axis=0)
    acc = e.copy()
    for _ in range(n_layers):
        e = adj @ e
        acc += e
    return acc / acc.dtype.type(n_layers + 1)


def lightgcn_loss_grads(p, adj, users, pos, neg, n_layers, decay):
    """LightGCNEngine.train_single_batch + loss_comput
    (beta_rec/models/lightgcn.py:119-152,171-191): softplus BPR on the
    propagated embeddings + decay * 0.5 * (|u0|^2+|p0|^2+|n0|^2)/B on layer-0 rows."""
    nu = p["user_embedding.weight"].shape[0]
    ni = p["item_embedding.weight"].shape[0]
    dt = p["user_embedding.weight"].dtype.type
    b = len(users)
    ebar = lightgcn_propagate(p, adj, n_layers)
    ue, ie = ebar[:nu], ebar[nu:]
    u, pi, nj = ue[users], ie[pos], ie[neg]
    ps = (u * pi).sum(1)
    ns = (u * nj).sum(1)
    mf_loss = softplus(ns - ps).mean()
    u0 = p["user_embedding.weight"][users]
    p0 = p["item_embedding.weight"][pos]
    n0 = p["item_embedding.weight"][neg]
    reg_loss = dt(0.5) * ((u0 ** 2).sum() + (p0 ** 2).sum() + (n0 ** 2).sum()) / dt(b) * dt(decay)
    # backward of the tail
    c = sigmoid(ns - ps) / dt(b)  # d mf_loss / d ns ; d/d ps = -c
    d_ebar = np.zeros_like(ebar)
    d_ebar[:nu] += _scatter_rows(users, c[:, None] * (nj - pi), nu)
    d_ebar[nu:] += _scatter_rows(pos, -c[:, None] * u, ni)
    d_ebar[nu:] += _scatter_rows(neg, c[:, None] * u, ni)
    # backward of the propagate: G_L = d/(L+1); G_l = d/(L+1) + Ahat^T G_{l+1}
    d = d_ebar / dt(n_layers + 1)
    gl = d.copy()
    at = adj.T.tocsr()
    for _ in range(n_layers):
        gl = d + at @ gl
    g_e0 = gl
    k = dt(decay) / dt(b)
    g_e0[:nu] += _scatter_rows(users, k * u0, nu)
    g_e0[nu:] += _scatter_rows(pos, k * p0, ni)
    g_e0[nu:] += _scatter_rows(neg, k * n0, ni)
    g = {"user_embedding.weight": g_e0[:nu].copy(), "item_embedding.weight": g_e0[nu:].copy()}
    return mf_loss + reg_loss, g


def lightgcn_train_single_batch(p, st, adj, users, pos, neg, n_layers, decay, optimizer="adam", lr=0.05):
    l, g = lightgcn_loss_grads(p, adj, users, pos, neg, n_layers, decay)
    optimizer_step(p, g, st, optimizer, lr)
    return float(l)


# --------------------------------------------------------------------------- #
# index / routing work (bit-exact domain)
# --------------------------------------------------------------------------- #
def owner_of(rows, world):
    """Row-sharding rule of the multi-GPU path: owner = row mod world, local
    row = row div world (new work -- the reference has no distributed path,
    SURVEY.md section 8e)."""
    rows = np.asarray(rows, dtype=np.int64)
    return rows % world, rows // world


def route_triples(users, pos, neg, world):
    """Stable bucket of (u,i,j) triples by owner(u): returns per-destination
    counts and the triples permuted so destination d's triples are contiguous,
    original order preserved inside a bucket."""
    own, _ = owner_of(users, world)
    order = np.argsort(own, kind="stable")
    counts = np.bincount(own, minlength=world).astype(np.int64)
    return counts, np.asarray(users)[order], np.asarray(pos)[order], np.asarray(neg)[order], order
