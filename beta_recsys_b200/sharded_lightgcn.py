"""Row-partitioned LightGCN on N GPUs of one node (BASELINE.json configs[3]: LightGCN 1M x 100k on 8 x B200).

The reference runs the whole-graph propagate of beta_rec/models/lightgcn.py:46-78 (L sparse products with the
[N, N] normalised adjacency, N = n_users + n_items) and its backward for EVERY batch; that is the step's cost.
Here the node rows are partitioned 1-D: rank r owns the r-th block of the USER rows and
the r-th block of the ITEM rows (two contiguous ranges; item rows are ~10x denser than user rows, so one contiguous
range per rank would leave a rank with most of the non-zeros) of the parameters E0 (+ optimizer state), of A_hat
and of A_hat^T (same pattern, so the backward is balanced too), and feeds its own batch.

  gather    E0 blocks -> every rank holds the full layer-0 matrix                 [NCCL all-gather x2: users, items]
  forward   per layer: E(l+1)[own rows] = A_hat[own rows, :] E(l)  (brs_spmm_csr per range, the edge-dropout mask
            folded in) then all-gather -> full E(l+1)
  tail      softplus-BPR + L2 on the rank's own batch, scaled for the GLOBAL batch mean (brs_lightgcn_tail):
            loss sum, sparse d = d loss / d E(l) into a full [N, D] buffer
  reduce    d summed over the ranks, each keeps its rows                          [NCCL reduce-scatter]
  backward  G(L) = d;  G(l)[own rows] = d[own rows] + A_hat^T[own rows, :] G(l+1)  -- all-gather between layers,
            none after the last: G(0)[own rows] is the gradient of the rank's own parameters
  regular.  + decay/B * E0[row] for the batch rows this rank owns (all batches are known everywhere after one
            small all-gather of the index arrays) (brs_lightgcn_reg_grad)
  step      Adam / SGD / RMSprop on the own rows (brs_dense_params_step); loss all-reduced.

Every rank draws the SAME edge-dropout mask (same seed, same generator state), like a single process would.
Per step and rank 2L + 1 collectives of N x D x 4 bytes (282 MB at config 4) replace (N-1)/N of the SpMM work.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engines.lightgcn import coo_to_csr
from .engines.torch_engine import RowOptimizer


def partition_rows(n_users, n_items, world, rank):
    """Host-side index rule (pure; covered by the CPU tests): the node rows rank `rank` owns -- block `rank` of the
    user rows and block `rank` of the item rows -- as a list of dicts {off, cnt, blk, lo, hi, own_off}: node offset
    and size of the entity, rows per rank (ceil), the own node range [lo, hi) and its offset inside the rank's own
    block (user part, then item part, each padded to blk rows)."""
    parts, own_off = [], 0
    for off, cnt in ((0, int(n_users)), (int(n_users), int(n_items))):
        blk = (cnt + world - 1) // world
        parts.append({"off": off, "cnt": cnt, "blk": blk, "lo": off + min(cnt, rank * blk), "hi": off + min(cnt, (rank + 1) * blk),
                      "own_off": own_off})
        own_off += blk
    return parts, own_off


class ShardedLightGCNEngine(object):
    def __init__(self, config, group=None, state=None):
        """config["model"]: the reference's LightGCN keys (lightgcn.py:104-117) with ``norm_adj`` either the reference's
        torch sparse tensor or a graph.GpuAdjacency; ``state``: full {"user_embedding.weight", "item_embedding.weight"}
        (numpy) or None for the reference's xavier_uniform init (identical on every rank)."""
        if not dist.is_initialized():
            raise _lib.BrsError("ShardedLightGCNEngine needs an initialised torch.distributed process group")
        m = config["model"]
        self.config, self.group = config, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(m["device_str"])
        if self.device.type != "cuda":
            raise _lib.BrsError("ShardedLightGCNEngine runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.load()
        dev, w = self.device, self.world
        self.n_users, self.n_items, self.dim = int(m["n_users"]), int(m["n_items"]), int(m["emb_dim"])
        self.n_layers = len(m["layer_size"])
        self.decay = float(m["regs"][0])
        self.keep_prob = float(m["keep_pro"])
        self.batch_size = int(m["batch_size"])
        self.opt_kind, self.lr = m["optimizer"], float(m["lr"])
        n = self.n_users + self.n_items
        self.n = n
        # ---- adjacency: full CSR arrays on the device (pattern of A_hat is symmetric: A_hat^T shares it)
        adj = m["norm_adj"]
        if hasattr(adj, "csr_tensors"):
            csr = {k: v.to(dev) for k, v in adj.csr_tensors().items()}
            self.nnz = adj.nnz
        else:
            a = adj.coalesce()
            idx = a.indices().cpu().numpy()
            c = coo_to_csr(idx[0], idx[1], a.values().cpu().numpy(), n)
            self.nnz = c["nnz"]
            csr = {k: torch.from_numpy(v).to(dev) for k, v in c.items() if k != "nnz"}
        self._csr = csr
        # ---- the rank's rows: block r of the user rows and block r of the item rows
        self.parts, self.own_rows = partition_rows(self.n_users, self.n_items, w, self.rank)
        for part in self.parts:
            part["fwd"] = self._block(csr["row_ptr"], csr["col"], csr["val"], None, part["lo"], part["hi"])
            part["bwd"] = self._block(csr["row_ptr_t"], csr["col_t"], csr["val_t"], csr["edge_id_t"], part["lo"], part["hi"])
            if part["cnt"] != part["blk"] * w:  # not divisible: collectives go through a padded staging copy of the entity's rows
                part["stage"] = torch.zeros((w * part["blk"], self.dim), dtype=torch.float32, device=dev)
        # ---- layer buffers (full), gradient chain buffers, the own block of the parameters
        f32 = torch.float32
        self._layers = [torch.zeros((n, self.dim), dtype=f32, device=dev) for _ in range(self.n_layers + 1)]
        self._d = torch.zeros((n, self.dim), dtype=f32, device=dev)
        self._g = [torch.zeros((n, self.dim), dtype=f32, device=dev) for _ in range(2)]
        self._d_blk = torch.zeros((self.own_rows, self.dim), dtype=f32, device=dev)
        self._tmp_blk = torch.zeros((self.own_rows, self.dim), dtype=f32, device=dev)
        self.param = torch.zeros((self.own_rows, self.dim), dtype=f32, device=dev)  # [user block | item block], zero padded
        self.grad = torch.zeros_like(self.param)
        if state is None:
            torch.manual_seed(2020)
            ue = torch.empty((self.n_users, self.dim))
            ie = torch.empty((self.n_items, self.dim))
            torch.nn.init.xavier_uniform_(ue)  # lightgcn.py:40-44
            torch.nn.init.xavier_uniform_(ie)
            full = torch.cat([ue, ie])
        else:
            full = torch.from_numpy(np.concatenate([np.asarray(state["user_embedding.weight"], dtype=np.float32),
                                                    np.asarray(state["item_embedding.weight"], dtype=np.float32)]))
        for pt in self.parts:
            self.param[pt["own_off"]: pt["own_off"] + pt["hi"] - pt["lo"]].copy_(full[pt["lo"]:pt["hi"]])
        self.opt = RowOptimizer(self.opt_kind, self.lr, "dense")
        self._st = self.opt.add_param("all_embeddings", self.param)
        self._t = 0
        self._ws = torch.zeros(_lib.STEP_WS_BYTES, dtype=torch.uint8, device=dev)
        # brs_lightgcn_model over the FULL buffers: used by the tail / regularizer entry points only
        c = _lib.LightGCNModel()
        c.n_users, c.n_items, c.dim, c.n_layers, c.decay = self.n_users, self.n_items, self.dim, self.n_layers, self.decay
        c.adj = _lib.Csr(_lib.ptr(csr["row_ptr"]), _lib.ptr(csr["col"]), _lib.ptr(csr["val"]), None, n, self.nnz)
        c.adj_t = _lib.Csr(_lib.ptr(csr["row_ptr_t"]), _lib.ptr(csr["col_t"]), _lib.ptr(csr["val_t"]), _lib.ptr(csr["edge_id_t"]),
                           n, self.nnz)
        for l, buf in enumerate(self._layers):
            c.emb[l] = _lib.ptr(buf)
        c.d = _lib.ptr(self._d)
        c.g[0], c.g[1] = _lib.ptr(self._g[0]), _lib.ptr(self._g[1])
        c.param = _lib.DenseParam(_lib.ptr(self._layers[0]), _lib.ptr(self._g[0]), None, None, n * self.dim)
        c.ws = _lib.ptr(self._ws)
        self._cmodel = c
        self._dense = _lib.DenseParam(_lib.ptr(self.param), _lib.ptr(self.grad), _lib.ptr(self._st.get("m")),
                                      _lib.ptr(self._st.get("v")), self.param.numel())
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(int(m["dropout_seed"]) if "dropout_seed" in m else 2020)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _block(self, row_ptr, col, val, edge_id, lo, hi):
        """CSR of the row range [lo, hi): rebased row_ptr, views into col / val / edge_id."""
        rp = row_ptr[lo:hi + 1].to(torch.int64) if hi > lo else torch.zeros(1, dtype=torch.int64, device=self.device)
        first = int(rp[0].item()) if hi > lo else 0
        last = int(rp[-1].item()) if hi > lo else 0
        local = (rp - first).to(torch.int32).contiguous()
        blk = {"row_ptr": local, "col": col[first:last], "val": val[first:last], "first": first, "nnz": last - first,
               "edge_id": None if edge_id is None else edge_id[first:last]}
        blk["struct"] = _lib.Csr(_lib.ptr(local), _lib.ptr(blk["col"]) if last > first else _lib.ptr(col),
                                 _lib.ptr(blk["val"]) if last > first else _lib.ptr(val),
                                 None if edge_id is None else (_lib.ptr(blk["edge_id"]) if last > first else _lib.ptr(edge_id)),
                                 max(hi - lo, 0), last - first)
        return blk

    def _spmm(self, blk, mask, x_full, y_blk):
        """y_blk += block x x_full with the keep mask (forward block: mask offset by the block's first edge)."""
        if blk["nnz"] == 0:
            return
        mp = None
        if mask is not None:
            mp = _lib.ptr(mask) if blk["edge_id"] is not None else mask.data_ptr() + blk["first"]
        _lib.check(self.lib.brs_spmm_csr(blk["struct"], mp, self.keep_prob, _lib.ptr(x_full), _lib.ptr(y_blk), self.dim,
                                         self._stream()), "brs_spmm_csr")

    def draw_keep_mask(self):
        """LightGCN.dropout's mask (lightgcn.py:32-33) on the device generator: identical on every rank."""
        return (torch.rand(self.nnz, device=self.device, generator=self._gen) + self.keep_prob).int().bool().to(torch.uint8)

    def _spmm_own(self, which, mask, x_full, y_own):
        """y_own[own rows] += (A_hat or A_hat^T)[own rows, :] x_full, one block product per range."""
        for pt in self.parts:
            self._spmm(pt[which], mask, x_full, y_own[pt["own_off"]: pt["own_off"] + pt["blk"]])

    def _gather_full(self, dst_full, own):
        """dst_full [n, D] <- every rank's block (own: [own_rows, D], user part then item part)."""
        for pt in self.parts:
            src = own[pt["own_off"]: pt["own_off"] + pt["blk"]]
            if "stage" in pt:
                dist.all_gather_into_tensor(pt["stage"], src.contiguous(), group=self.group)
                dst_full[pt["off"]: pt["off"] + pt["cnt"]].copy_(pt["stage"][: pt["cnt"]])
            else:
                dist.all_gather_into_tensor(dst_full[pt["off"]: pt["off"] + pt["cnt"]], src.contiguous(), group=self.group)

    def _reduce_own(self, src_full, out_own):
        """out_own <- sum over ranks of src_full[own rows]."""
        for pt in self.parts:
            dst = out_own[pt["own_off"]: pt["own_off"] + pt["blk"]]
            if "stage" in pt:
                pt["stage"][: pt["cnt"]].copy_(src_full[pt["off"]: pt["off"] + pt["cnt"]])
                pt["stage"][pt["cnt"]:].zero_()
                dist.reduce_scatter_tensor(dst, pt["stage"], op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.reduce_scatter_tensor(dst, src_full[pt["off"]: pt["off"] + pt["cnt"]], op=dist.ReduceOp.SUM, group=self.group)

    def propagate(self, keep_mask):
        self._gather_full(self._layers[0], self.param)
        for l in range(self.n_layers):
            own = self._tmp_blk
            own.zero_()
            self._spmm_own("fwd", keep_mask, self._layers[l], own)
            self._gather_full(self._layers[l + 1], own)

    def train_single_batch(self, batch_data, keep_mask=None):
        """LightGCNEngine.train_single_batch (lightgcn.py:119-152) on this rank's batch; returns the GLOBAL batch loss."""
        dev, lib, st = self.device, self.lib, self._stream
        users, pos, neg = (torch.as_tensor(t).to(dev, torch.int64).contiguous().view(-1) for t in batch_data)
        b = users.numel()
        if b != self.batch_size or pos.numel() != b or neg.numel() != b:
            raise ValueError("every rank feeds exactly batch_size samples per step")
        gb = b * self.world
        bad = ((users < 0) | (users >= self.n_users) | (pos < 0) | (pos >= self.n_items) | (neg < 0) | (neg >= self.n_items)).any()
        flag = bad.to(torch.float32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        if float(flag.item()) != 0.0:
            raise IndexError("index out of range in self")
        if keep_mask is None:
            keep_mask = self.draw_keep_mask()
        else:
            keep_mask = torch.as_tensor(keep_mask).to(torch.uint8).to(dev).contiguous()
            if keep_mask.numel() != self.nnz:
                raise ValueError("keep_mask must have one entry per coalesced edge of norm_adj")
        self.propagate(keep_mask)
        # ---- tail on the own batch (mean over the global batch); d summed over ranks, own rows kept
        _lib.check(lib.brs_lightgcn_tail(self._cmodel, _lib.ptr(users), _lib.ptr(pos), _lib.ptr(neg), b, gb, st()), "brs_lightgcn_tail")
        self._reduce_own(self._d, self._d_blk)
        # ---- backward of the propagate: G(L) = d, G(l) = d + A_hat^T G(l+1)
        ci = 1
        self._gather_full(self._g[ci], self._d_blk)
        for l in range(self.n_layers - 1, -1, -1):
            if l == 0:
                self.grad.copy_(self._d_blk)
                self._spmm_own("bwd", keep_mask, self._g[ci], self.grad)
            else:
                own = self._tmp_blk
                own.copy_(self._d_blk)
                self._spmm_own("bwd", keep_mask, self._g[ci], own)
                self._gather_full(self._g[1 - ci], own)
                ci = 1 - ci
        # ---- L2 term on the layer-0 rows of EVERY rank's batch that fall into the own rows
        mine = torch.stack([users, pos, neg])
        everyone = torch.empty((self.world,) + tuple(mine.shape), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        for pt in self.parts:
            if pt["hi"] <= pt["lo"]:
                continue
            gpart = self.grad[pt["own_off"]:]
            for r in range(self.world):
                u, p, q = everyone[r]
                _lib.check(lib.brs_lightgcn_reg_grad(self._cmodel, _lib.ptr(u), _lib.ptr(p), _lib.ptr(q), b, gb, pt["lo"], pt["hi"],
                                                     _lib.ptr(gpart), st()), "brs_lightgcn_reg_grad")
        # ---- optimizer on the own rows, loss over the global batch
        self._t += 1
        _lib.check(lib.brs_dense_params_step(self._dense, 1, self.opt.desc, self._t, st()), "brs_dense_params_step")
        loss = self._ws[0:8].view(torch.float64).clone()
        self._ws[0:8].zero_()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return float(loss.item()) / gb

    def gather_state(self):
        """{"user_embedding.weight", "item_embedding.weight"} (numpy), identical on every rank."""
        full = torch.empty((self.n, self.dim), dtype=torch.float32, device=self.device)
        self._gather_full(full, self.param)
        full = full.cpu().numpy()
        return {"user_embedding.weight": full[: self.n_users].copy(), "item_embedding.weight": full[self.n_users:].copy()}
