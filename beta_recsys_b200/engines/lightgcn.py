"""LightGCN module + engine: drop-in for beta_rec.models.lightgcn (same constructor,
methods and ``state_dict`` keys); the per-batch work runs in libbrs_b200
(csrc/lightgcn_kernels.cu: CSR propagate, softplus-BPR tail, transposed backward)."""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from .rows import as_index
from .torch_engine import ModelEngine


def coo_to_csr(rows, cols, vals, n):
    """Coalesced (row-major sorted) COO -> CSR arrays + the CSR of the transpose with the
    map from its non-zeros back to the forward edge ids (host side, once per engine)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float32)
    nnz = rows.shape[0]
    if nnz >= 2 ** 31 or n >= 2 ** 31:
        raise _lib.BrsError("adjacency too large for int32 CSR")
    key = rows * n + cols
    if nnz and np.any(np.diff(key) <= 0):  # the reference coalesces first (lightgcn.py:59)
        order = np.argsort(key, kind="stable")
        rows, cols, vals, key = rows[order], cols[order], vals[order], key[order]
        uniq, start = np.unique(key, return_index=True)
        vals = np.add.reduceat(vals, start).astype(np.float32)
        rows, cols = rows[start], cols[start]
        nnz = rows.shape[0]
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, rows + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    order_t = np.lexsort((rows, cols))  # sort by (col, row): the transpose's CSR order
    row_ptr_t = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr_t, cols + 1, 1)
    row_ptr_t = np.cumsum(row_ptr_t)
    return {
        "row_ptr": row_ptr.astype(np.int32), "col": cols.astype(np.int32), "val": vals,
        "row_ptr_t": row_ptr_t.astype(np.int32), "col_t": rows[order_t].astype(np.int32), "val_t": vals[order_t],
        "edge_id_t": order_t.astype(np.int32), "nnz": int(nnz),
    }


class LightGCN(nn.Module):
    """Parameters of beta_rec.models.lightgcn.LightGCN (lightgcn.py:10-44)."""

    def __init__(self, config, norm_adj):
        super(LightGCN, self).__init__()
        self.config = config
        self.n_users = config["n_users"]
        self.n_items = config["n_items"]
        self.emb_dim = config["emb_dim"]
        self.layer_size = config["layer_size"]
        self.n_layers = len(self.layer_size)
        self.norm_adj = norm_adj
        self.layer_size = [self.emb_dim] + self.layer_size
        self.user_embedding = nn.Embedding(self.n_users, self.emb_dim)
        self.item_embedding = nn.Embedding(self.n_items, self.emb_dim)
        self.f = nn.Sigmoid()
        self.init_emb()
        for p in self.parameters():
            p.requires_grad_(False)
        self._engine = None

    def init_emb(self):
        nn.init.xavier_uniform_(self.user_embedding.weight)
        nn.init.xavier_uniform_(self.item_embedding.weight)

    def forward(self, norm_adj=None):
        """Propagated (layer-mean) user and item embeddings (lightgcn.py:46-78).  In training
        mode the reference's edge dropout is applied with a mask drawn the reference's way."""
        eng = self._engine
        eng.propagate(eng.draw_keep_mask() if self.training else None)
        ebar = sum(eng._layers) / float(self.n_layers + 1)
        return ebar[: self.n_users], ebar[self.n_users:]

    def predict(self, users, items):
        """lightgcn.py:80-101: eval mode, full propagate without dropout, sigmoid(u.i)."""
        self.eval()
        users_t = torch.as_tensor(np.asarray(users), dtype=torch.int64).to(self.device)
        items_t = torch.as_tensor(np.asarray(items), dtype=torch.int64).to(self.device)
        return self._engine.scores(users_t, items_t)


class LightGCNEngine(ModelEngine):
    """Drop-in for beta_rec.models.lightgcn.LightGCNEngine (lightgcn.py:104-191).

    ``config["model"]["dropout_rng"]`` (new, optional): ``"cpu"`` (default) draws the edge
    keep mask exactly like LightGCN.dropout -- ``torch.rand(nnz)`` on the global CPU
    generator -- so runs are bit-identical in the mask to the reference; ``"cuda"`` draws it
    on the device (fast, different random stream)."""

    def __init__(self, config):
        self.config = config
        self.regs = config["model"]["regs"]
        self.decay = self.regs[0]
        self.norm_adj = config["model"]["norm_adj"]
        self.model = LightGCN(config["model"], self.norm_adj)
        super(LightGCNEngine, self).__init__(config)
        self.model.to(self.device)
        self.keep_prob = float(config["model"]["keep_pro"])
        self.dropout_rng = config["model"]["dropout_rng"] if "dropout_rng" in config["model"] else "cpu"
        self._bind()

    def _bind(self):
        m, dev = self.model, self.device
        n, d = m.n_users + m.n_items, m.emb_dim
        # one contiguous [N, D] parameter block (torch.cat at lightgcn.py:55-57 without the copy);
        # the two nn.Embedding weights become views of it, state_dict keys/shapes unchanged
        self._all = torch.empty((n, d), dtype=torch.float32, device=dev)
        self._all[: m.n_users].copy_(m.user_embedding.weight.data)
        self._all[m.n_users:].copy_(m.item_embedding.weight.data)
        m.user_embedding.weight.data = self._all[: m.n_users]
        m.item_embedding.weight.data = self._all[m.n_users:]
        if hasattr(self.norm_adj, "csr_tensors"):  # graph.GpuAdjacency: built on the device, nothing to convert
            if self.norm_adj.n != n:
                raise ValueError("norm_adj is %d x %d, the model has %d nodes" % (self.norm_adj.n, self.norm_adj.n, n))
            self._nnz = self.norm_adj.nnz
            self._csr = {k: v.to(dev) for k, v in self.norm_adj.csr_tensors().items()}
        else:  # the reference's torch sparse tensor (recommenders/lightgcn.py:15-23)
            adj = self.norm_adj.coalesce()
            idx = adj.indices().cpu().numpy()
            csr = coo_to_csr(idx[0], idx[1], adj.values().cpu().numpy(), n)
            self._nnz = csr["nnz"]
            self._csr = {k: torch.from_numpy(v).to(dev) for k, v in csr.items() if k != "nnz"}
        self._layers = [self._all] + [torch.zeros((n, d), dtype=torch.float32, device=dev) for _ in range(m.n_layers)]
        self._d = torch.zeros((n, d), dtype=torch.float32, device=dev)
        self._g = [torch.zeros((n, d), dtype=torch.float32, device=dev) for _ in range(2)]
        self._grad = torch.zeros((n, d), dtype=torch.float32, device=dev)
        self._st = self.optimizer.add_param("all_embeddings", self._all)
        self._ws = torch.zeros(_lib.STEP_WS_BYTES, dtype=torch.uint8, device=dev)
        self._out = torch.zeros(4, dtype=torch.float32, device=dev)
        c = _lib.LightGCNModel()
        c.n_users, c.n_items, c.dim, c.n_layers, c.decay = m.n_users, m.n_items, d, m.n_layers, float(self.decay)
        t = self._csr
        c.adj = _lib.Csr(_lib.ptr(t["row_ptr"]), _lib.ptr(t["col"]), _lib.ptr(t["val"]), None, n, self._nnz)
        c.adj_t = _lib.Csr(_lib.ptr(t["row_ptr_t"]), _lib.ptr(t["col_t"]), _lib.ptr(t["val_t"]),
                           _lib.ptr(t["edge_id_t"]), n, self._nnz)
        for l, buf in enumerate(self._layers):
            c.emb[l] = _lib.ptr(buf)
        c.d = _lib.ptr(self._d)
        c.g[0], c.g[1] = _lib.ptr(self._g[0]), _lib.ptr(self._g[1])
        c.param = _lib.DenseParam(_lib.ptr(self._all), _lib.ptr(self._grad), _lib.ptr(self._st.get("m")),
                                  _lib.ptr(self._st.get("v")), self._all.numel())
        c.ws = _lib.ptr(self._ws)
        self._cmodel = c
        m._engine = self

    # ------------------------------------------------------------------ #
    def draw_keep_mask(self):
        """LightGCN.dropout's mask (lightgcn.py:32-33): (rand(nnz) + keep_prob).int().bool()."""
        if self.dropout_rng == "cpu":
            return (torch.rand(self._nnz) + self.keep_prob).int().bool().to(torch.uint8).to(self.device)
        return (torch.rand(self._nnz, device=self.device) + self.keep_prob).int().bool().to(torch.uint8)

    def propagate(self, keep_mask):
        lib = _lib.load()
        _lib.check(lib.brs_lightgcn_propagate(self._cmodel, _lib.ptr(keep_mask), self.keep_prob, self._stream()),
                   "brs_lightgcn_propagate")

    def scores(self, users, items):
        lib = _lib.load()
        users, items = as_index(users, self.device), as_index(items, self.device)
        self.propagate(None)
        out = torch.full((users.numel(),), float("nan"), dtype=torch.float32, device=self.device)
        _lib.check(lib.brs_lightgcn_scores(self._cmodel, _lib.ptr(users), _lib.ptr(items), users.numel(), _lib.ptr(out),
                                           self._stream()), "brs_lightgcn_scores")
        self._check_predict()
        return out

    def train_single_batch(self, batch_data, keep_mask=None):
        """lightgcn.py:119-152: whole-graph propagate with edge dropout, BPR tail, backward, optimizer."""
        assert hasattr(self, "model"), "Please specify the exact model !"
        lib = _lib.load()
        users, pos, neg = (as_index(t, self.device) for t in batch_data)
        b = users.numel()
        if keep_mask is None and self.model.training:
            keep_mask = self.draw_keep_mask()
        elif keep_mask is not None:
            keep_mask = torch.as_tensor(keep_mask).to(torch.uint8).to(self.device).contiguous()
            if keep_mask.numel() != self._nnz:
                raise ValueError("keep_mask must have one entry per coalesced edge of norm_adj")
        _lib.check(lib.brs_lightgcn_fwd_bwd(self._cmodel, _lib.ptr(keep_mask), self.keep_prob, _lib.ptr(users),
                                            _lib.ptr(pos), _lib.ptr(neg), b, self._stream()), "brs_lightgcn_fwd_bwd")
        _lib.check(lib.brs_lightgcn_apply(self._cmodel, self.optimizer.desc, b, _lib.ptr(self._out), self._stream()),
                   "brs_lightgcn_apply")
        loss, _, status = _lib.step_record(self._out)
        if int(status) & 1:
            raise IndexError("index out of range in self")
        return loss

    def train_an_epoch(self, train_loader, epoch_id):
        """lightgcn.py:154-169."""
        assert hasattr(self, "model"), "Please specify the exact model !"
        self.model.train()
        total_loss, loss = 0.0, 0.0
        for batch_data in train_loader:
            loss = self.train_single_batch(batch_data)
            total_loss += loss
        print("[Training Epoch {}], Loss {}".format(epoch_id, loss))
        self.writer.add_scalar("model/loss", total_loss, epoch_id)
