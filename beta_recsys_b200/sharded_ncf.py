"""Row-sharded NeuMF on N GPUs of one node (BASELINE.json configs[2]: NeuMF 10M x 1M on 8 x B200).

The four embedding tables of beta_rec/models/ncf.py:40-47 are row-sharded (owner = row mod N, like the MF
tables of sharded.py); the MLP tower and the output layer (fc_layers.*, affine_output: ~170 k parameters) are
replicated.  One step, every rank feeding its own batch of (user, item, rating):

  route     ids bucketed by owner -> NCCL all-to-all (variable splits)          [torch.distributed]
  serve     every owner gathers the requested rows of its shards (brs_gather)   [csrc/abi.cu rows_op_kernel]
  return    rows back to the requesting rank -> the batch's activations          [NCCL all-to-all]
  compute   NeuMF forward + BCE + backward on the rank's own batch: the stock single-GPU kernels
            (brs_ncf_fwd_bwd: gather / tcgen05 tower / head / scatter) run on a per-batch "identity" table
            whose row s is sample s's embedding row, so row s's gradient is d loss / d (sample s's row)
  reduce    Linear-layer gradients averaged over the ranks (NCCL all-reduce), identical Adam / SGD step on
            every replica (brs_ncf_apply)
  push      per-sample gradient rows (brs_rows_read_grad) -> owners (NCCL all-to-all) -> summed per row into
            the owner's compact scratch (brs_rows_assign + brs_rows_scatter_grad, scale 1/N: the loss is the mean
            over the GLOBAL batch) -> row optimizer on the owner (brs_rows_sgd / brs_rows_adam /
            brs_dense_adam_sweep)

This is the exchange pattern north_star names (one all-to-all to the owning rank) with rows instead of triples:
the tower needs all four rows of a sample on one rank.  Per rank and step 2 x B x (mlp_dim + emb_dim) x 4 bytes
cross NVLink in each direction twice (168 MB at B = 65 536, emb_dim 64, 3 layers).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engines.ncf import NeuMFEngine
from .engines.rows import EntityState
from .engines.torch_engine import RowOptimizer
from .sharded import all_to_all_v, local_rows, shard_of, unshard

_TABLES = (("user", "embedding_user_mlp.weight", "mlp"), ("user", "embedding_user_mf.weight", "mf"),
           ("item", "embedding_item_mlp.weight", "mlp"), ("item", "embedding_item_mf.weight", "mf"))


def bucket_by_owner(ids, world):
    """Host-side index rule of the exchange (pure; covered by the CPU tests): the permutation that sorts a batch's
    ids by owner = id mod world (stable), the number of ids per owner, and the owners' local row numbers in that
    order.  Works on CPU and CUDA tensors."""
    owner = ids % world
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)
    return order, counts, (ids // world)[order].contiguous()


class ShardedNeuMFEngine(object):
    def __init__(self, config, group=None, state=None):
        """config["model"]: the reference's NeuMF keys (ncf.py:82-98) + adam_mode; ``state``: a full (unsharded)
        state dict with the reference's keys (numpy), or None for the reference's own initialisation."""
        if not dist.is_initialized():
            raise _lib.BrsError("ShardedNeuMFEngine needs an initialised torch.distributed process group")
        m = config["model"]
        self.config, self.group = config, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(m["device_str"])
        if self.device.type != "cuda":
            raise _lib.BrsError("ShardedNeuMFEngine runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.load()
        self.n_users, self.n_items = int(m["n_users"]), int(m["n_items"])
        self.emb, self.n_layers = int(m["emb_dim"]), int(m["mlp_config"]["n_layers"])
        self.mlp_dim = self.emb * 2 ** (self.n_layers - 1)
        self.batch_size = int(m["batch_size"])
        self.opt_kind, self.lr = m["optimizer"], float(m["lr"])
        self.mode = m["adam_mode"] if "adam_mode" in m else "dense"
        dev, w = self.device, self.world
        # ---- home side: the stock engine over per-batch identity tables (row s = sample s) + the replicated tower
        home_cfg = {"model": dict(m, n_users=self.batch_size, n_items=self.batch_size, adam_mode="touched"),
                    "system": config["system"] if "system" in config else {"run_dir": None}}
        torch.manual_seed(2020)  # identical tower on every rank (also broadcast below)
        self.home = NeuMFEngine(home_cfg)
        self._dense_names = [n for n, _, _, _ in self.home._dense]
        with torch.no_grad():
            for _, wt, _, _ in self.home._dense:
                dist.broadcast(wt, src=0, group=group)
        self._ident = torch.arange(self.batch_size, dtype=torch.int64, device=dev)
        # ---- owner side: the shards + their optimizer state and compact gradient scratch
        self.opt = RowOptimizer(self.opt_kind, self.lr, self.mode)
        lu, li = local_rows(self.n_users, w), local_rows(self.n_items, w)
        self.local_users, self.local_items = lu, li
        g = torch.Generator(device=dev)
        g.manual_seed(2020 + self.rank)
        self.shards = {}
        for ent, key, kind in _TABLES:
            rows, dim = (lu if ent == "user" else li), (self.mlp_dim if kind == "mlp" else self.emb)
            t = torch.empty((rows, dim), dtype=torch.float32, device=dev)
            # ncf.py:142-154: user tables and the item MF table N(0, 0.01^2); the item MLP table keeps nn.Embedding's N(0, 1)
            t.normal_(0, 1.0 if key == "embedding_item_mlp.weight" else 0.01, generator=g)
            if state is not None:
                t.copy_(torch.from_numpy(shard_of(np.asarray(state[key], dtype=np.float32), w, self.rank)))
            self.shards[key] = t
        if state is not None:
            with torch.no_grad():
                for name, wt, _, _ in self.home._dense:
                    wt.copy_(torch.from_numpy(np.asarray(state[name], dtype=np.float32)).view_as(wt))
        cap = w * self.batch_size
        self._user = EntityState(lu, [(k, self.shards[k]) for e, k, _ in _TABLES if e == "user"], self.opt, cap, dev)
        self._item = EntityState(li, [(k, self.shards[k]) for e, k, _ in _TABLES if e == "item"], self.opt, cap, dev)
        self._ws = torch.zeros(_lib.STEP_WS_BYTES, dtype=torch.uint8, device=dev)
        self._t = 0
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---------------------------------------------------------------- exchange
    def _route(self, ids):
        """Bucket the batch's ids by owner and send each owner its local row numbers.  Returns (order, send_counts,
        recv_counts, recv_local_rows): ``order`` sorts the batch by owner (stable)."""
        order, counts, local = bucket_by_owner(ids, self.world)
        send_counts = counts.tolist()
        recv, recv_counts = all_to_all_v(local, send_counts, self.group)
        return order, send_counts, recv_counts, recv

    def _exchange(self, send, in_counts, out_counts):
        recv = send.new_empty((int(sum(out_counts)),) + tuple(send.shape[1:]))
        dist.all_to_all_single(recv, send.contiguous(), list(out_counts), list(in_counts), group=self.group)
        return recv

    def _serve(self, table, rows):
        out = torch.empty((rows.numel(), table.shape[1]), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.brs_gather(_lib.ptr(table), table.shape[0], table.shape[1], _lib.ptr(rows), rows.numel(),
                                       _lib.ptr(out), self._stream()), "brs_gather")
        return out

    # ---------------------------------------------------------------- step
    def train_single_batch(self, users, items, ratings):
        """NeuMFEngine.train_single_batch (ncf.py:100-120) on this rank's batch; returns the GLOBAL batch loss."""
        dev, lib, st = self.device, self.lib, self._stream
        users = torch.as_tensor(users).to(dev, torch.int64).contiguous().view(-1)
        items = torch.as_tensor(items).to(dev, torch.int64).contiguous().view(-1)
        ratings = torch.as_tensor(ratings).to(dev, torch.float32).contiguous().view(-1)
        b = users.numel()
        if b != self.batch_size or items.numel() != b or ratings.numel() != b:
            raise ValueError("every rank feeds exactly batch_size samples per step")
        bad = ((users < 0) | (users >= self.n_users) | (items < 0) | (items >= self.n_items)).any()
        flag = bad.to(torch.float32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        if float(flag.item()) != 0.0:  # the reference raises IndexError inside nn.Embedding, before any update
            raise IndexError("index out of range in self")
        hm = self.home.model
        routes = {}
        for ent, ids, home_tabs in (("user", users, (hm.embedding_user_mlp, hm.embedding_user_mf)),
                                    ("item", items, (hm.embedding_item_mlp, hm.embedding_item_mf))):
            order, sc, rc, rows = self._route(ids)
            routes[ent] = (order, sc, rc, rows)
            keys = [k for e, k, _ in _TABLES if e == ent]
            for key, tab in zip(keys, home_tabs):
                back = self._exchange(self._serve(self.shards[key], rows), rc, sc)  # sorted by owner
                tab.weight.data.index_copy_(0, order, back)  # row s of the identity table = sample s's row
        # ---- NeuMF forward + BCE + backward on the identity tables (mean over the LOCAL batch)
        hc = self.home._cmodel
        _lib.check(lib.brs_ncf_fwd_bwd(hc, _lib.ptr(self._ident), _lib.ptr(self._ident), _lib.ptr(ratings), b, st()),
                   "brs_ncf_fwd_bwd")
        grads = {}
        for ent, es in (("user", self.home._user), ("item", self.home._item)):
            for tno, key in enumerate(k for e, k, _ in _TABLES if e == ent):
                g = torch.empty((b, self.shards[key].shape[1]), dtype=torch.float32, device=dev)
                _lib.check(lib.brs_rows_read_grad(es.struct, tno, _lib.ptr(self._ident), b, _lib.ptr(g), st()), "brs_rows_read_grad")
                grads[key] = g
        for _, _, gbuf, _ in self.home._dense:  # replicated tower: average the Linear gradients over the ranks
            dist.all_reduce(gbuf, op=dist.ReduceOp.AVG, group=self.group)
        _lib.check(lib.brs_ncf_apply(hc, self.home.optimizer.desc, b, _lib.ptr(self.home._out), st()), "brs_ncf_apply")
        # ---- per-sample gradient rows -> owners, summed per row, row optimizer on the owner
        self._t += 1
        ents = []
        for ent, es in (("user", self._user), ("item", self._item)):
            order, sc, rc, rows = routes[ent]
            _lib.check(lib.brs_rows_assign(es.struct.rows, _lib.ptr(rows), rows.numel(), _lib.ptr(self._ws), st()), "brs_rows_assign")
            for tno, key in enumerate(k for e, k, _ in _TABLES if e == ent):
                recv = self._exchange(grads[key][order], sc, rc)
                _lib.check(lib.brs_rows_scatter_grad(es.struct, tno, _lib.ptr(rows), rows.numel(), _lib.ptr(recv),
                                                     1.0 / self.world, st()), "brs_rows_scatter_grad")
            ents.append(es.struct)
        arr = (_lib.Entity * 2)(*ents)
        if self.opt_kind == "sgd":
            _lib.check(lib.brs_rows_sgd(arr, 2, self.lr, st()), "brs_rows_sgd")
        elif self.mode == "touched" and self.opt_kind == "adam":
            _lib.check(lib.brs_rows_adam(arr, 2, self.opt.desc, self._t, st()), "brs_rows_adam")
        else:  # reference-exact: every row moves (g = 0 outside the batch)
            _lib.check(lib.brs_dense_adam_sweep(arr, 2, self.opt.desc, self._t, st()), "brs_dense_adam_sweep")
        loss = self.home._out[:1].clone()
        dist.all_reduce(loss, op=dist.ReduceOp.AVG, group=self.group)
        return float(loss.item())

    # ---------------------------------------------------------------- state
    def gather_state(self):
        """Full state dict in the reference layout (numpy), identical on every rank."""
        out = {}
        for ent, key, _ in _TABLES:
            n = self.n_users if ent == "user" else self.n_items
            local = self.shards[key]
            parts = [torch.empty_like(local) for _ in range(self.world)]
            dist.all_gather(parts, local.contiguous(), group=self.group)
            out[key] = unshard([p.cpu().numpy() for p in parts], n)
        for name, wt, _, _ in self.home._dense:
            out[name] = wt.detach().cpu().numpy().copy()
        return out

    def save_checkpoint(self, model_dir):
        """torch_engine.py:70-73: the reference module's state_dict, written by rank 0 (collective)."""
        state = self.gather_state()
        if self.rank == 0:
            torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in state.items()}, model_dir)
        dist.barrier(group=self.group)
