"""Row-sharded multi-GPU MF engine (one process per GPU, torch.distributed for plumbing).

New work -- the reference has no distributed path.  Tables are
row-sharded: ``owner(row) = row mod world``, ``local row = row div world``.  Every rank
exports its shard (weights; for Adam / RMSprop also a dense per-shard gradient table and a touched
bitmap) through CUDA IPC; the kernels address remote memory directly over NVLink.  A step
(``brs_mf_sharded_step``) is

    [optional NCCL all-to-all: route triples to the user-row owner]
    local slot pre-pass: every unique row of the rank's batch gets a compact local slot
    rows: either gathered per sample with peer loads inside the fused kernel ("direct"), or every
          UNIQUE row pulled once into local staging tables and the fused kernel run on those
          ("staged", default from 8 ranks); gradients are summed in a LOCAL compact scratch
          (no remote atomics per sample)
    push: one coalesced row of 128-bit peer REDs per unique touched row --
          SGD: -lr * g straight into the owner's weight rows (flag barrier first: all gathers done);
          Adam / RMSprop: g into the owner's dense gradient table + red.or of its touched bit,
          flag barrier, optimizer on the rows of the local shard whose bit is set
    global_bias is replicated and updated identically on every rank (the first barrier also
          exchanges the 3 step sums, added in rank order); flag barrier

``train_batches`` runs many steps per C call, from index arrays in HBM or (CPU tensors) streamed
from pinned host memory through a device ring.

The loss is the mean over the GLOBAL batch (sum of the ranks' batches), exactly what a
single-GPU run on the concatenated batch computes.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engines.rows import as_index


# --------------------------------------------------------------------------- #
# host-side index rules (pure functions; covered by the CPU tests)
# --------------------------------------------------------------------------- #
def owner_of(rows, world):
    """(owner rank, local row) of global row ids under the interleaved sharding."""
    return rows % world, rows // world


def local_rows(n_rows, world):
    """Rows every rank reserves for a table of n_rows global rows (uniform, ceil)."""
    return (int(n_rows) + world - 1) // world


def shard_of(full, world, rank):
    """Rows of a full [N, ...] array owned by `rank`, padded to local_rows(N, world)."""
    part = full[rank::world]
    need = local_rows(full.shape[0], world)
    if part.shape[0] < need:
        pad = np.zeros((need - part.shape[0],) + full.shape[1:], dtype=full.dtype)
        part = np.concatenate([part, pad], axis=0)
    return np.ascontiguousarray(part)


def unshard(parts, n_rows):
    """Inverse of shard_of over all ranks: parts[r] is rank r's [local_rows, ...] array."""
    world = len(parts)
    out = np.zeros((n_rows,) + parts[0].shape[1:], dtype=parts[0].dtype)
    for r, p in enumerate(parts):
        cnt = (n_rows - r + world - 1) // world
        out[r::world] = p[:cnt]
    return out


def all_to_all_v(send, send_counts, group=None):
    """Variable-size all-to-all of a 1-D/2-D tensor whose dim 0 is grouped by destination.
    NCCL: one all_to_all_single; other backends (gloo in the CPU tests): point-to-point."""
    world = dist.get_world_size(group)
    sc = torch.as_tensor(send_counts, dtype=torch.int64)
    rc = torch.empty(world, dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        sc_d, rc_d = sc.to(send.device), torch.empty(world, dtype=torch.int64, device=send.device)
        dist.all_to_all_single(rc_d, sc_d, group=group)
        rc = rc_d.cpu()
    else:
        gathered = [torch.empty(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, sc, group=group)
        me = dist.get_rank(group)
        rc = torch.stack([g[me] for g in gathered])
    recv = send.new_empty((int(rc.sum()),) + tuple(send.shape[1:]))
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv, send, rc.tolist(), sc.tolist(), group=group)
    else:
        me = dist.get_rank(group)
        so = np.concatenate([[0], np.cumsum(sc.numpy())])
        ro = np.concatenate([[0], np.cumsum(rc.numpy())])
        reqs = []
        for r in range(world):
            if r == me:
                recv[ro[r]:ro[r + 1]] = send[so[r]:so[r + 1]]
                continue
            reqs.append(dist.isend(send[so[r]:so[r + 1]].contiguous(), r, group=group))
            reqs.append(dist.irecv(recv[ro[r]:ro[r + 1]], r, group=group))
        for q in reqs:
            q.wait()
    return recv, rc.tolist()


# --------------------------------------------------------------------------- #
# exportable device memory
# --------------------------------------------------------------------------- #
class _Raw(object):
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class PeerArena(object):
    """One cudaMalloc'ed, IPC-exported block per rank with the SAME named layout on every
    rank, so that (peer base + offset[name]) addresses any peer's buffer."""

    ALIGN = 256

    def __init__(self, layout, device, group=None):
        lib = _lib.load()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = device
        self.offsets, self.specs, off = {}, {}, 0
        for name, shape, dtype in layout:
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            self.offsets[name] = off
            self.specs[name] = (tuple(int(x) for x in shape), dtype, nbytes)
            off += (nbytes + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.nbytes = max(off, self.ALIGN)
        base = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.brs_shm_alloc(self.nbytes, C.byref(base)), "brs_shm_alloc")
            self.base = base.value
            self._raw = torch.as_tensor(_Raw(self.base, self.nbytes), device=device)
            handle = (C.c_uint8 * _lib.IPC_HANDLE_BYTES)()
            _lib.check(lib.brs_ipc_get_handle(self.base, handle), "brs_ipc_get_handle")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self.peer_base = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.peer_base.append(self.base)
                    continue
                buf = (C.c_uint8 * _lib.IPC_HANDLE_BYTES).from_buffer_copy(h)
                p = C.c_void_p()
                _lib.check(lib.brs_ipc_open_handle(buf, C.byref(p)), "brs_ipc_open_handle")
                self.peer_base.append(p.value)

    def tensor(self, name):
        shape, dtype, nbytes = self.specs[name]
        o = self.offsets[name]
        return self._raw[o:o + nbytes].view(dtype).view(shape)

    def ptr(self, name, rank=None):
        return (self.peer_base[self.rank if rank is None else rank]) + self.offsets[name]

    def close(self):
        lib = _lib.load()
        dist.barrier(group=self.group)  # nobody may still be reading our memory
        for r, p in enumerate(self.peer_base):
            if r != self.rank and p:
                lib.brs_ipc_close_handle(p)
        self.peer_base = []
        self._raw = None
        if self.base:
            lib.brs_shm_free(self.base)
            self.base = None


# --------------------------------------------------------------------------- #
# the engine
# --------------------------------------------------------------------------- #
class ShardedMFEngine(object):
    """MF-BPR over row-sharded tables.  ``config["model"]`` as for MFEngine (GLOBAL
    n_users / n_items; batch_size is PER RANK); ``route`` = "none" (default: every rank
    trains its own triples, all rows through peer memory) or "owner" (NCCL all-to-all
    routes each triple to the rank that owns its user row first)."""

    def __init__(self, config, group=None, route="none", state=None):
        if not dist.is_initialized():
            raise _lib.BrsError("ShardedMFEngine needs an initialised torch.distributed process group")
        m = config["model"]
        self.config, self.group, self.route = config, group, route
        if route not in ("none", "owner"):
            raise ValueError("route must be 'none' or 'owner'")
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.MAX_RANKS or self.world & (self.world - 1):
            raise _lib.BrsError("world size must be a power of two <= %d" % _lib.MAX_RANKS)
        self.device = torch.device(m["device_str"])
        if self.device.type != "cuda":
            raise _lib.BrsError("ShardedMFEngine runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.load()
        self.n_users, self.n_items, self.dim = int(m["n_users"]), int(m["n_items"]), int(m["emb_dim"])
        self.batch_size = int(m["batch_size"])
        self.reg = config["model"]["reg"] if "reg" in config else 0.0  # mf.py:81-83 quirk
        if (m["loss"] if "loss" in m else "bpr") != "bpr":
            raise _lib.BrsError("ShardedMFEngine implements the BPR loss")
        self.opt_kind, self.lr = m["optimizer"], float(m["lr"])
        mode = m["adam_mode"] if "adam_mode" in m else "dense"
        self.opt = _lib.make_opt(self.opt_kind, self.lr, _lib.DENSE if mode == "dense" else _lib.TOUCHED_ROWS)
        w, d = self.world, self.dim
        lu, li = local_rows(self.n_users, w), local_rows(self.n_items, w)
        self.local_users, self.local_items = lu, li
        f32, i32 = torch.float32, torch.int32
        layout = [  # peer-visible: shard weights, dense shard gradients, touched bitmaps, barrier state
            ("user_emb", (lu, d), f32), ("item_emb", (li, d), f32), ("user_bias", (lu, 1), f32),
            ("item_bias", (li, 1), f32), ("g_user_emb", (lu, d), f32), ("g_item_emb", (li, d), f32),
            ("g_user_bias", (lu,), f32), ("g_item_bias", (li,), f32), ("user_bits", ((lu + 31) // 32,), i32),
            ("item_bits", ((li + 31) // 32,), i32), ("flags", (_lib.MAX_RANKS,), torch.int64),
            ("partials", (_lib.MAX_RANKS * 4,), torch.float64), ("ws", (_lib.STEP_WS_BYTES,), torch.uint8),
        ]
        self.arena = PeerArena(layout, self.device, group)
        t = self.arena.tensor
        # local staging (not peer-visible): slot maps over GLOBAL ids + compact gradient scratch
        grow = 2 if route == "owner" else 1  # routed batches are uneven (Zipf users)
        self.cap_u = int(max(1, min(self.n_users, grow * self.batch_size)))
        self.cap_i = int(max(1, min(self.n_items, 2 * grow * self.batch_size)))
        dev = self.device
        self._stage = {
            "user_slot": torch.full((self.n_users,), -1, dtype=i32, device=dev),
            "item_slot": torch.full((self.n_items,), -1, dtype=i32, device=dev),
            "user_list": torch.zeros(self.cap_u, dtype=i32, device=dev),
            "item_list": torch.zeros(self.cap_i, dtype=i32, device=dev),
            "user_count": torch.zeros(1, dtype=i32, device=dev), "item_count": torch.zeros(1, dtype=i32, device=dev),
            # staging tables of the pulled unique rows (one row per slot)
            "p_user_emb": torch.zeros(self.cap_u * d, dtype=f32, device=dev),
            "p_item_emb": torch.zeros(self.cap_i * d, dtype=f32, device=dev),
            "p_user_bias": torch.zeros(self.cap_u, dtype=f32, device=dev),
            "p_item_bias": torch.zeros(self.cap_i, dtype=f32, device=dev),
            "s_user_emb": torch.zeros(self.cap_u * d, dtype=f32, device=dev),
            "s_item_emb": torch.zeros(self.cap_i * d, dtype=f32, device=dev),
            "s_user_bias": torch.zeros(self.cap_u, dtype=f32, device=dev),
            "s_item_bias": torch.zeros(self.cap_i, dtype=f32, device=dev),
        }
        self.global_bias = torch.zeros(1, dtype=f32, device=self.device)  # replicated
        self._init_tables(state)
        # optimizer state: local only
        self.state = {}
        for name in ("user_emb", "item_emb", "user_bias", "item_bias"):
            self.state[name] = self._new_state(t(name))
        self.state["global_bias"] = self._new_state(self.global_bias)
        self._out = torch.zeros(4, dtype=f32, device=self.device)
        self._epoch = 0
        self._build_structs()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    # -- construction helpers ------------------------------------------------ #
    def _new_state(self, tensor):
        st = {}
        if self.opt_kind == "adam":
            st["m"] = torch.zeros_like(tensor)
        if self.opt_kind in ("adam", "rmsprop"):
            st["v"] = torch.zeros_like(tensor)
        return st

    def _init_tables(self, state):
        """state: full (unsharded) numpy state dict with the reference keys, or None for
        MF's own init (N(0, 0.1^2) embeddings, zero biases; mf.py:25-30) seeded per rank."""
        t = self.arena.tensor
        if state is not None:
            for name, key in (("user_emb", "user_emb.weight"), ("item_emb", "item_emb.weight"),
                              ("user_bias", "user_bias.weight"), ("item_bias", "item_bias.weight")):
                t(name).copy_(torch.from_numpy(shard_of(np.asarray(state[key], dtype=np.float32), self.world, self.rank)))
            self.global_bias.copy_(torch.from_numpy(np.asarray(state["global_bias"], dtype=np.float32)))
        else:
            g = torch.Generator(device=self.device)
            g.manual_seed(2020 + self.rank)
            t("user_emb").normal_(0, 0.1, generator=g)
            t("item_emb").normal_(0, 0.1, generator=g)

    def _build_structs(self):
        A, S = self.arena, self._stage
        w, d = self.world, self.dim
        lu, li = self.local_users, self.local_items

        def table(name, scratch, rows, dim):
            st = self.state[name]
            return _lib.Table(A.ptr(name), _lib.ptr(S[scratch]), _lib.ptr(st.get("m")), _lib.ptr(st.get("v")), rows, dim, 0)

        def entity(prefix, n_global, rows, cap):
            e = _lib.Entity()
            e.rows = _lib.Rowset(_lib.ptr(S[prefix + "_slot"]), _lib.ptr(S[prefix + "_list"]),
                                 _lib.ptr(S[prefix + "_count"]), n_global, cap, 0)  # slot maps over GLOBAL ids
            e.n_tables = 2
            e.table[0] = table(prefix + "_emb", "s_" + prefix + "_emb", rows, d)
            e.table[1] = table(prefix + "_bias", "s_" + prefix + "_bias", rows, 1)
            return e

        gb = self.state["global_bias"]
        stage = _lib.MfModel(entity("user", self.n_users, lu, self.cap_u), entity("item", self.n_items, li, self.cap_i),
                             _lib.DenseParam(_lib.ptr(self.global_bias), None, _lib.ptr(gb.get("m")),
                                             _lib.ptr(gb.get("v")), 1), A.ptr("ws"), _lib.Rowset(), _lib.Rowset())
        peers = (_lib.MfPeerTables * w)()
        for r in range(w):
            for f, _ in _lib.MfPeerTables._fields_:
                setattr(peers[r], f, A.ptr(f, r))
        self._peers_dev = torch.frombuffer(bytearray(bytes(peers)), dtype=torch.uint8).to(self.device)
        own = _lib.MfPeerTables()
        for f, _ in _lib.MfPeerTables._fields_:
            setattr(own, f, A.ptr(f))
        self._cmodel = _lib.MfSharded(w, self.rank, self.n_users, self.n_items, lu, li, stage,
                                      self._peers_dev.data_ptr(), own, _lib.ptr(S["p_user_emb"]),
                                      _lib.ptr(S["p_item_emb"]), _lib.ptr(S["p_user_bias"]), _lib.ptr(S["p_item_bias"]))
        self._sync = _lib.PeerSync()
        self._sync.world, self._sync.rank = w, self.rank
        for r in range(w):
            self._sync.flags[r] = A.ptr("flags", r)
            self._sync.partials[r] = A.ptr("partials", r)

    # -- step ---------------------------------------------------------------- #
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def route_triples(self, users, pos, neg):
        """NCCL all-to-all of triples to the rank owning the user row (north-star routing)."""
        n = users.numel()
        ou, op_, on = torch.empty_like(users), torch.empty_like(pos), torch.empty_like(neg)
        counts = torch.empty(self.world, dtype=torch.int64, device=self.device)
        _lib.check(self.lib.brs_route_triples(_lib.ptr(users), _lib.ptr(pos), _lib.ptr(neg), n, self.world,
                                              _lib.ptr(ou), _lib.ptr(op_), _lib.ptr(on), _lib.ptr(counts),
                                              self._stream()), "brs_route_triples")
        send = torch.stack([ou, op_, on], dim=1)  # [n, 3] int64, grouped by destination
        recv, _ = all_to_all_v(send, counts.cpu().tolist(), self.group)
        return recv[:, 0].contiguous(), recv[:, 1].contiguous(), recv[:, 2].contiguous()

    def launch_step(self, batch, global_batch=None, out=None):
        """Enqueue one step (no host sync unless routing is on).  `batch`: this rank's
        (users, pos, neg) GLOBAL ids.  global_batch defaults to world * len(batch)."""
        users, pos, neg = (as_index(x, self.device) for x in batch)
        gb = int(global_batch) if global_batch is not None else users.numel() * self.world
        if self.route == "owner":
            users, pos, neg = self.route_triples(users, pos, neg)
        _lib.check(self.lib.brs_mf_sharded_step(
            C.byref(self._cmodel), C.byref(self._sync), C.byref(self.opt), _lib.ptr(users), _lib.ptr(pos), _lib.ptr(neg),
            users.numel(), gb, float(self.reg), self._epoch + 1, _lib.ptr(self._out if out is None else out),
            self._stream()), "brs_mf_sharded_step")
        self._epoch += 2

    def train_batches(self, users, pos, neg):
        """Many consecutive steps over this rank's index arrays (route='none'): one C call, no host
        involvement until the records are read.  Device tensors: arrays resident in HBM, returns a device
        tensor [n_batches, 4].  CPU tensors (pin them): the C loop streams batch b+2 through a device ring
        while batch b computes and DMAs every step's record back; returns a numpy array."""
        if self.route != "none":
            raise _lib.BrsError("train_batches needs route='none' (routing needs a host round trip per batch)")
        host = all(isinstance(t, torch.Tensor) and not t.is_cuda for t in (users, pos, neg))
        if host:
            users, pos, neg = (t.to(torch.int64).contiguous() for t in (users, pos, neg))
        else:
            users, pos, neg = (as_index(x, self.device) for x in (users, pos, neg))
        n, b = users.numel(), self.batch_size
        if pos.numel() != n or neg.numel() != n:
            raise ValueError("users / pos / neg must have the same length")
        n_batches = (n + b - 1) // b
        if host:
            import numpy as np

            out = np.zeros((n_batches, 4), dtype=np.float32)
            with torch.cuda.device(self.device):
                _lib.check(self.lib.brs_mf_sharded_train_batches_host(
                    C.byref(self._cmodel), C.byref(self._sync), C.byref(self.opt), users.data_ptr(), pos.data_ptr(),
                    neg.data_ptr(), n, b, b * self.world, float(self.reg), self._epoch + 1, out.ctypes.data,
                    self._stream()), "brs_mf_sharded_train_batches_host")
            self._epoch += 2 * n_batches
            return out
        out = torch.zeros((n_batches, 4), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.brs_mf_sharded_train_batches(
            C.byref(self._cmodel), C.byref(self._sync), C.byref(self.opt), _lib.ptr(users), _lib.ptr(pos), _lib.ptr(neg),
            n, b, b * self.world, float(self.reg), self._epoch + 1, _lib.ptr(out), self._stream()),
            "brs_mf_sharded_train_batches")
        self._epoch += 2 * n_batches
        return out

    def train_single_batch(self, batch, global_batch=None):
        self.launch_step(batch, global_batch)
        loss, reg, status = _lib.step_record(self._out)
        if int(status) & 1:
            raise IndexError("index out of range in self")
        if status:
            raise _lib.BrsError("touched-row capacity overflow")
        return loss, reg

    # -- state --------------------------------------------------------------- #
    def gather_state(self):
        """Full state dict in the reference layout (numpy), identical on every rank."""
        t = self.arena.tensor
        out = {"global_bias": self.global_bias.cpu().numpy().copy()}
        for name, key, n in (("user_emb", "user_emb.weight", self.n_users), ("item_emb", "item_emb.weight", self.n_items),
                             ("user_bias", "user_bias.weight", self.n_users), ("item_bias", "item_bias.weight", self.n_items)):
            parts = [torch.empty_like(t(name)) for _ in range(self.world)]
            dist.all_gather(parts, t(name).contiguous(), group=self.group)
            out[key] = unshard([p.cpu().numpy() for p in parts], n)
        return out

    _TABLES = (("user_emb", "user_emb.weight", "n_users"), ("item_emb", "item_emb.weight", "n_items"),
               ("user_bias", "user_bias.weight", "n_users"), ("item_bias", "item_bias.weight", "n_items"))

    def _gather_full(self, local, n_rows):
        parts = [torch.empty_like(local) for _ in range(self.world)]
        dist.all_gather(parts, local.contiguous(), group=self.group)
        return unshard([p.cpu().numpy() for p in parts], n_rows)

    def gather_optimizer_state(self):
        """Optimizer state in the layout of torch.optim's state_dict()["state"] for the reference module's
        parameter order (global_bias, user_emb, item_emb, user_bias, item_bias -- models/mf.py:17-30):
        {index: {"step", "exp_avg", "exp_avg_sq"}} for Adam, {"step", "square_avg"} for RMSprop, {} for SGD.
        Identical on every rank."""
        step = int(self.arena.tensor("ws")[24:32].view(torch.int64).item())  # brs_step_ws.step
        names = {"m": "exp_avg", "v": "exp_avg_sq"} if self.opt_kind == "adam" else {"v": "square_avg"}
        out = {}
        order = [("global_bias", None, None)] + [(n, k, r) for n, k, r in self._TABLES]
        for idx, (name, _, rows_attr) in enumerate(order):
            st = self.state[name]
            if not st:
                continue
            ent = {"step": torch.tensor(float(step))}
            for kind, tname in names.items():
                if kind not in st:
                    continue
                full = st[kind].cpu().numpy().copy() if rows_attr is None else self._gather_full(st[kind], getattr(self, rows_attr))
                ent[tname] = torch.from_numpy(full)
            out[idx] = ent
        return out

    def save_checkpoint(self, model_dir):
        """ModelEngine.save_checkpoint (beta_rec/models/torch_engine.py:70-73) for the sharded tables: the shards
        are gathered into the reference module's ``state_dict`` (same keys, shapes, dtypes) and rank 0 writes it
        with ``torch.save`` -- the file loads into the reference's own ``MF`` module or into the single-GPU
        ``MFEngine``.  The optimizer state goes to ``model_dir + ".optim"`` (the reference does not save it;
        resuming Adam without it restarts the moments).  Collective: every rank must call it."""
        state = self.gather_state()
        opt_state = self.gather_optimizer_state()
        if self.rank == 0:
            torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in state.items()}, model_dir)
            torch.save({"optimizer": self.opt_kind, "lr": self.lr, "state": opt_state}, model_dir + ".optim")
        dist.barrier(group=self.group)

    def resume_checkpoint(self, model_dir, load_optimizer=True):
        """ModelEngine.resume_checkpoint (torch_engine.py:76-90): every rank reads the reference-layout file and
        keeps its own rows (owner = row mod world); the replicated global bias is copied whole.  Works for a
        checkpoint written by the reference, by MFEngine or by any world size of this engine."""
        print("loading model from:", model_dir)
        sd = torch.load(model_dir, map_location="cpu")
        want = {"global_bias"} | {k for _, k, _ in self._TABLES}
        if set(sd) != want:
            raise RuntimeError("checkpoint keys do not match the model: %s" % sorted(set(sd) ^ want))
        t = self.arena.tensor
        with torch.no_grad():
            for name, key, rows_attr in self._TABLES:
                full = sd[key].numpy()
                if full.shape[0] != getattr(self, rows_attr) or full.shape[1:] != tuple(t(name).shape[1:]):
                    raise RuntimeError("size mismatch for %s: checkpoint %s" % (key, tuple(full.shape)))
                t(name).copy_(torch.from_numpy(shard_of(full.astype(np.float32, copy=False), self.world, self.rank)))
            self.global_bias.copy_(sd["global_bias"].to(torch.float32).view(-1))
            opt_path = model_dir + ".optim"
            import os

            if load_optimizer and self.opt_kind != "sgd" and os.path.exists(opt_path):
                od = torch.load(opt_path, map_location="cpu")
                if od["optimizer"] != self.opt_kind:
                    raise RuntimeError("optimizer state is for %s, the engine runs %s" % (od["optimizer"], self.opt_kind))
                names = {"m": "exp_avg", "v": "exp_avg_sq"} if self.opt_kind == "adam" else {"v": "square_avg"}
                order = [("global_bias", None)] + [(n, r) for n, _, r in self._TABLES]
                step = 0
                for idx, (name, rows_attr) in enumerate(order):
                    ent = od["state"].get(idx)
                    if ent is None:
                        continue
                    step = int(ent["step"])
                    for kind, tname in names.items():
                        full = ent[tname].numpy().astype(np.float32, copy=False)
                        part = full if rows_attr is None else shard_of(full, self.world, self.rank)
                        self.state[name][kind].copy_(torch.from_numpy(part).view_as(self.state[name][kind]))
                self.arena.tensor("ws")[24:32].view(torch.int64).fill_(step)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def close(self):
        torch.cuda.synchronize(self.device)
        self.arena.close()
