"""MF module + MFEngine: drop-in for beta_rec.models.mf (same constructor, same
methods, same ``state_dict`` keys) with the per-batch work done by the fused
sm_100a kernels of libbrs_b200 (csrc/mf_kernels.cu, csrc/rows_apply.cu)."""
import time

import numpy as np
import torch
import torch.nn as nn
from torch.nn import Parameter

from .. import _lib
from .rows import EntityState, as_float, as_index, dense_param, loader_index_batches
from .torch_engine import ModelEngine


class MF(nn.Module):
    """Parameters of beta_rec.models.mf.MF (mf.py:12-30): user_emb, item_emb,
    user_bias, item_bias, global_bias -- same names, shapes and initialisation."""

    def __init__(self, config):
        super(MF, self).__init__()
        self.config = config
        self.device = self.config["device_str"]
        self.stddev = self.config["stddev"] if "stddev" in self.config else 0.1
        self.n_users = self.config["n_users"]
        self.n_items = self.config["n_items"]
        self.emb_dim = self.config["emb_dim"]
        self.user_emb = nn.Embedding(self.n_users, self.emb_dim)
        self.item_emb = nn.Embedding(self.n_items, self.emb_dim)
        self.user_bias = nn.Embedding(self.n_users, 1)
        self.item_bias = nn.Embedding(self.n_items, 1)
        self.global_bias = Parameter(torch.zeros(1))
        self.user_bias.weight.data.fill_(0.0)
        self.item_bias.weight.data.fill_(0.0)
        self.global_bias.data.fill_(0.0)
        nn.init.normal_(self.user_emb.weight, 0, self.stddev)
        nn.init.normal_(self.item_emb.weight, 0, self.stddev)
        for p in self.parameters():  # gradients are produced by the kernels, never by autograd
            p.requires_grad_(False)
        self._engine = None

    def forward(self, batch_data):
        """(users, items) LongTensors -> (sigmoid scores, regularizer) as mf.py:32-55.
        Scores come from the predict kernel; the regularizer (only reported by the
        reference, weight 0) is evaluated with the same formula on the gathered rows."""
        users, items = batch_data
        scores = self._engine.scores(users, items)
        with torch.no_grad():
            u, i = self.user_emb.weight[users], self.item_emb.weight[items]
            ub, ib = self.user_bias.weight[users], self.item_bias.weight[items]
            regularizer = ((u ** 2).sum() + (i ** 2).sum() + (ub ** 2).sum() + (ib ** 2).sum()) / u.size()[0]
        return scores, regularizer

    def predict(self, users, items):
        """mf.py:57-70: numpy / list ids -> scores tensor on the model's device."""
        users_t = torch.as_tensor(np.asarray(users), dtype=torch.int64).to(self.device)
        items_t = torch.as_tensor(np.asarray(items), dtype=torch.int64).to(self.device)
        return self._engine.scores(users_t, items_t)


DEFAULT_STEP_IMPL = "rows"


class MFEngine(ModelEngine):
    """Drop-in for beta_rec.models.mf.MFEngine (mf.py:73-139)."""

    def __init__(self, config):
        self.config = config
        self.model = MF(config["model"])
        # mf.py:81-83 tests the TOP-LEVEL config, so this is 0.0 for every stock config
        self.reg = config["model"]["reg"] if "reg" in config else 0.0
        self.batch_size = config["model"]["batch_size"]
        super(MFEngine, self).__init__(config)
        self.model.to(self.device)
        self.loss = self.config["model"]["loss"] if "loss" in self.config["model"] else "bpr"
        print(f"using {self.loss} loss...")
        self._bind()

    # ------------------------------------------------------------------ #
    def _bind(self):
        """Allocate the kernel scratch and freeze raw pointers into brs_mf_model."""
        m, dev, opt = self.model, self.device, self.optimizer
        for p in m.parameters():
            assert p.is_cuda and p.is_contiguous()
        b = int(self.batch_size)
        self._user = EntityState(m.n_users, [("user_emb.weight", m.user_emb.weight.data),
                                             ("user_bias.weight", m.user_bias.weight.data)], opt, b, dev)
        self._item = EntityState(m.n_items, [("item_emb.weight", m.item_emb.weight.data),
                                             ("item_bias.weight", m.item_bias.weight.data)], opt, 2 * b, dev)
        self._gb_state = opt.add_param("global_bias", m.global_bias.data)
        self._ws = torch.zeros(_lib.STEP_WS_BYTES, dtype=torch.uint8, device=dev)
        self._out = torch.zeros(4, dtype=torch.float32, device=dev)
        # "rows" (default): the row-owner step of csrc/mf_rowwise.cu; "scratch": round 1's fused gather/loss/RED
        # kernel + gradient-scratch apply (csrc/mf_kernels.cu, rows_apply.cu), kept as a second implementation
        mc = self.config["model"]
        self._step_impl = mc["step_impl"] if "step_impl" in mc else DEFAULT_STEP_IMPL
        if self._step_impl not in ("rows", "scratch"):
            raise ValueError("step_impl must be 'rows' or 'scratch'")
        self._plan_batch = 0
        self._refresh_struct(b)
        m._engine = self

    def _refresh_struct(self, b=None):
        self._cmodel = _lib.MfModel(self._user.struct, self._item.struct,
                                    dense_param(self.model.global_bias.data, None, self._gb_state),
                                    _lib.ptr(self._ws), self._user.alt_rowset(), self._item.alt_rowset())
        if self._step_impl != "rows":
            return
        lib, dev = _lib.load(), self.device
        self._plan_batch = max(int(b or 0), self._plan_batch, 1)
        nbytes = lib.brs_mf_plan_bytes(self._plan_batch, self._user.capacity, self._item.capacity)
        self._plans = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
        self._user_stage = torch.empty((self._user.capacity, self.model.emb_dim), dtype=torch.float32, device=dev)
        for k in range(2):
            self._cmodel.plan[k] = _lib.MfPlan(_lib.ptr(self._plans[k]), nbytes, self._plan_batch, self._user.capacity,
                                               self._item.capacity)
        self._cmodel.user_stage = _lib.ptr(self._user_stage)

    def _ensure_capacity(self, b):
        grew = self._user.ensure_capacity(b)
        grew = self._item.ensure_capacity(2 * b) or grew
        if grew or (self._step_impl == "rows" and b > self._plan_batch):
            self._refresh_struct(b)

    @staticmethod
    def _raise_status(status):
        if status == 0:
            return
        if int(status) & 1:
            raise IndexError("index out of range in self")  # what nn.Embedding raises in the reference
        raise _lib.BrsError("touched-row list overflow (internal capacity error)")

    # ------------------------------------------------------------------ #
    def scores(self, users, items):
        lib = _lib.load()
        users, items = as_index(users, self.device), as_index(items, self.device)
        if users.numel() != items.numel():
            raise ValueError("users and items must have the same length")
        out = torch.empty(users.numel(), dtype=torch.float32, device=self.device)
        _lib.check(lib.brs_mf_predict(self._cmodel, _lib.ptr(users), _lib.ptr(items), users.numel(), _lib.ptr(out),
                                      self._stream()), "brs_mf_predict")
        self._check_predict()
        return out

    def _launch_step(self, batch_data, out):
        lib = _lib.load()
        if self.loss not in ("bpr", "bce"):
            raise RuntimeError(f"Unsupported loss type {self.loss}, try other options: 'bpr' or 'bce'")
        if self._step_impl == "rows":
            users, items, third = batch_data
            users, items = as_index(users, self.device), as_index(items, self.device)
            third = as_index(third, self.device) if self.loss == "bpr" else as_float(third, self.device)
            b = users.numel()
            if items.numel() != b or third.numel() != b:
                raise ValueError("users / items / third must have the same length")
            self._ensure_capacity(b)
            _lib.check(lib.brs_mf_step(self._cmodel, self.optimizer.desc, 0 if self.loss == "bpr" else 1,
                                       _lib.ptr(users), _lib.ptr(items), _lib.ptr(third), b, float(self.reg),
                                       _lib.ptr(out), self._stream()), "brs_mf_step")
            return
        if self.loss == "bpr":
            users, pos, neg = batch_data
            users, pos, neg = (as_index(t, self.device) for t in (users, pos, neg))
            b = users.numel()
            if pos.numel() != b or neg.numel() != b:
                raise ValueError("users / pos_items / neg_items must have the same length")
            self._ensure_capacity(b)
            _lib.check(lib.brs_mf_bpr_fwd_bwd(self._cmodel, _lib.ptr(users), _lib.ptr(pos), _lib.ptr(neg), b,
                                              float(self.reg), self._stream()), "brs_mf_bpr_fwd_bwd")
        elif self.loss == "bce":
            users, items, ratings = batch_data
            users, items = as_index(users, self.device), as_index(items, self.device)
            ratings = as_float(ratings, self.device)
            b = users.numel()
            if items.numel() != b or ratings.numel() != b:
                raise ValueError("users / items / ratings must have the same length")
            self._ensure_capacity(b)
            _lib.check(lib.brs_mf_bce_fwd_bwd(self._cmodel, _lib.ptr(users), _lib.ptr(items), _lib.ptr(ratings), b,
                                              float(self.reg), self._stream()), "brs_mf_bce_fwd_bwd")
        else:
            raise RuntimeError(f"Unsupported loss type {self.loss}, try other options: 'bpr' or 'bce'")
        _lib.check(lib.brs_mf_apply(self._cmodel, self.optimizer.desc, b, _lib.ptr(out), self._stream()),
                   "brs_mf_apply")

    def train_single_batch(self, batch_data):
        """mf.py:92-119: one batch -> (loss: float, regularizer: float); one C call + one 16-byte D2H."""
        assert hasattr(self, "model"), "Please specify the exact model !"
        self._launch_step(batch_data, self._out)
        loss, reg, status = _lib.step_record(self._out)  # the reference's two .item() syncs, in one copy
        self._raise_status(status)
        return loss, reg

    def train_batches(self, users, items, third):
        """Inner loop of train_an_epoch over index arrays resident in HBM (one C call,
        2 launches per batch, no host sync until the per-batch results are read)."""
        lib = _lib.load()
        if all(isinstance(t, torch.Tensor) and not t.is_cuda for t in (users, items, third)):
            return self._train_batches_host(users, items, third)
        users, items = as_index(users, self.device), as_index(items, self.device)
        third = as_index(third, self.device) if self.loss == "bpr" else as_float(third, self.device)
        n, b = users.numel(), int(self.batch_size)
        if n == 0:
            return np.zeros((0, 4), dtype=np.float32)
        self._ensure_capacity(min(n, b))
        n_batches = (n + b - 1) // b
        out = torch.zeros((n_batches, 4), dtype=torch.float32, device=self.device)
        _lib.check(lib.brs_mf_train_batches(self._cmodel, self.optimizer.desc, 0 if self.loss == "bpr" else 1,
                                            _lib.ptr(users), _lib.ptr(items), _lib.ptr(third), n, b, float(self.reg),
                                            _lib.ptr(out), self._stream()), "brs_mf_train_batches")
        res = out.cpu().numpy()
        self._raise_status(_lib.step_records_status(res))
        return res

    def _train_batches_host(self, users, items, third):
        """train_batches for HOST tensors (pin them for full speed): the C loop streams batch b+2 to the
        device while batch b computes and brings every step's record back; nothing is staged in bulk."""
        lib = _lib.load()
        users = users.to(torch.int64).contiguous()
        items = items.to(torch.int64).contiguous()
        third = (third.to(torch.int64) if self.loss == "bpr" else third.to(torch.float32)).contiguous()
        n, b = users.numel(), int(self.batch_size)
        if items.numel() != n or third.numel() != n:
            raise ValueError("users / items / third must have the same length")
        if n == 0:
            return np.zeros((0, 4), dtype=np.float32)
        self._ensure_capacity(min(n, b))
        res = np.zeros(((n + b - 1) // b, 4), dtype=np.float32)
        with torch.cuda.device(self.device):
            _lib.check(lib.brs_mf_train_batches_host(self._cmodel, self.optimizer.desc, 0 if self.loss == "bpr" else 1,
                                                     users.data_ptr(), items.data_ptr(), third.data_ptr(), n, b,
                                                     float(self.reg), res.ctypes.data, self._stream()),
                       "brs_mf_train_batches_host")
        self._raise_status(_lib.step_records_status(res))
        return res

    def train_an_epoch(self, train_loader, epoch_id):
        """mf.py:121-139.  When the loader is the reference's DataLoader over a
        Pairwise/Rating dataset whose tensors already live on the device
        (base_data.py:247-253), the shuffled batches are gathered on the device
        in the loader's own order and trained by one C call; otherwise batches are
        taken from the loader one by one like the reference does."""
        assert hasattr(self, "model"), "Please specify the exact model !"
        t0 = time.time()
        self.model.train()
        total_loss, regularizer, loss = 0.0, 0.0, 0.0
        fast = self._epoch_tensors(train_loader)
        if fast is not None:
            res = self.train_batches(*fast)
            for l, r in zip(res[:, 0].tolist(), res[:, 1].tolist()):
                loss = l
                total_loss += l
                regularizer += r
        else:
            for batch_data in train_loader:
                loss, reg = self.train_single_batch(batch_data)
                total_loss += loss
                regularizer += reg
        print(f"[Training Epoch {epoch_id}], Loss {loss}, Regularizer {regularizer}")
        self.writer.add_scalar("model/loss", total_loss, epoch_id)
        self.writer.add_scalar("model/regularizer", regularizer, epoch_id)
        print("Execute [train_an_epoch] method costing %.2f ms" % ((time.time() - t0) * 1000))

    def _epoch_tensors(self, loader):
        ds = getattr(loader, "dataset", None)
        if self.loss == "bpr":
            names = ("user_tensor", "pos_item_tensor", "neg_item_tensor")  # data_loaders.py:30-53
        else:
            names = ("user_tensor", "item_tensor", "target_tensor")  # data_loaders.py:4-27
        if ds is None or not all(hasattr(ds, n) for n in names):
            return None
        if getattr(loader, "drop_last", False) or getattr(loader, "batch_size", None) != self.batch_size:
            return None
        batches = loader_index_batches(loader)
        if batches is None:
            return None
        order = torch.as_tensor([i for b in batches for i in b], dtype=torch.int64).to(self.device)
        cols = [getattr(ds, n).to(self.device) for n in names]
        return tuple(c.index_select(0, order) for c in cols)
