"""GMF / MLP / NeuMF modules + engines: drop-ins for beta_rec.models.{gmf,mlp,ncf}
(same constructors, methods and ``state_dict`` keys); the per-batch work runs in
libbrs_b200 (csrc/ncf_kernels.cu, csrc/gemm_*.cu, csrc/rows_apply.cu)."""
import os
import time

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from .rows import EntityState, as_float, as_index, dense_param, loader_index_batches
from .torch_engine import ModelEngine


def _freeze(module):
    for p in module.parameters():
        p.requires_grad_(False)


class _NcfModule(nn.Module):
    """Shared predict/forward plumbing: scores come from the forward kernels."""

    def forward(self, user_indices, item_indices):
        return self._engine.scores(user_indices, item_indices).view(-1, 1)  # the reference returns [B,1]

    def predict(self, user_indices, item_indices):
        user_indices = torch.as_tensor(np.asarray(user_indices), dtype=torch.int64).to(self.device)
        item_indices = torch.as_tensor(np.asarray(item_indices), dtype=torch.int64).to(self.device)
        with torch.no_grad():
            return self.forward(user_indices, item_indices)


class GMF(_NcfModule):
    """Parameters of beta_rec.models.gmf.GMF (gmf.py:11-27)."""

    def __init__(self, config):
        super(GMF, self).__init__()
        self.config = config
        self.num_users = config["n_users"]
        self.num_items = config["n_items"]
        self.emb_dim = config["emb_dim"]
        self.embedding_user = nn.Embedding(num_embeddings=self.num_users, embedding_dim=self.emb_dim)
        self.embedding_item = nn.Embedding(num_embeddings=self.num_items, embedding_dim=self.emb_dim)
        self.init_weight()
        self.affine_output = nn.Linear(in_features=self.emb_dim, out_features=1)
        self.logistic = nn.Sigmoid()
        _freeze(self)
        self._engine = None

    def init_weight(self):
        """gmf.py:45-48 initialises the USER table twice (item table keeps N(0,1))."""
        nn.init.normal_(self.embedding_user.weight, std=0.01)
        nn.init.normal_(self.embedding_user.weight, std=0.01)


def _fc_layers(emb_dim, n_layers, dropout):
    mods = []
    for i in range(n_layers):
        input_size = emb_dim * (2 ** (n_layers - i))
        mods.append(nn.Dropout(p=dropout))
        mods.append(nn.Linear(input_size, input_size // 2))
        mods.append(nn.ReLU())
    return nn.Sequential(*mods)


class MLP(_NcfModule):
    """Parameters of beta_rec.models.mlp.MLP (mlp.py:11-38)."""

    def __init__(self, config):
        super(MLP, self).__init__()
        self.config = config
        self.n_users = config["n_users"]
        self.n_items = config["n_items"]
        self.emb_dim = config["emb_dim"]
        self.n_layers = config["mlp_config"]["n_layers"]
        self.dropout = config["dropout"]
        self.latent_dim = self.emb_dim * (2 ** (self.n_layers)) // 2
        self.embedding_user = nn.Embedding(num_embeddings=self.n_users, embedding_dim=self.latent_dim)
        self.embedding_item = nn.Embedding(num_embeddings=self.n_items, embedding_dim=self.latent_dim)
        self.init_weight()
        self.fc_layers = _fc_layers(self.emb_dim, self.n_layers, self.dropout)
        self.affine_output = nn.Linear(in_features=self.emb_dim, out_features=1)
        self.logistic = nn.Sigmoid()
        _freeze(self)
        self._engine = None

    def init_weight(self):
        nn.init.normal_(self.embedding_user.weight, std=0.01)
        nn.init.normal_(self.embedding_user.weight, std=0.01)


class NeuMF(_NcfModule):
    """Parameters of beta_rec.models.ncf.NeuMF (ncf.py:15-50)."""

    def __init__(self, config):
        super(NeuMF, self).__init__()
        self.config = config
        self.n_users = config["n_users"]
        self.n_items = config["n_items"]
        self.emb_dim = config["emb_dim"]
        self.n_layers = config["mlp_config"]["n_layers"]
        self.dropout = config["dropout"]
        self.latent_dim_mlp = self.emb_dim * (2 ** (self.n_layers)) // 2
        self.latent_dim_gmf = self.emb_dim
        self.embedding_user_mlp = nn.Embedding(num_embeddings=self.n_users, embedding_dim=self.latent_dim_mlp)
        self.embedding_item_mlp = nn.Embedding(num_embeddings=self.n_items, embedding_dim=self.latent_dim_mlp)
        self.embedding_user_mf = nn.Embedding(num_embeddings=self.n_users, embedding_dim=self.latent_dim_gmf)
        self.embedding_item_mf = nn.Embedding(num_embeddings=self.n_items, embedding_dim=self.latent_dim_gmf)
        self.fc_layers = _fc_layers(self.emb_dim, self.n_layers, self.dropout)
        self.affine_output = nn.Linear(in_features=self.emb_dim * 2, out_features=1)
        self.logistic = nn.Sigmoid()
        _freeze(self)
        self._engine = None

    def init_weight(self):
        pass


class _NcfEngine(ModelEngine):
    """Common engine: binds the module's tensors into brs_ncf_model and drives the kernels."""

    KIND = None

    def _common_init(self, config):
        if float(config["model"]["dropout"] if "dropout" in config["model"] else 0.0) != 0.0 and self.KIND != _lib.NCF_GMF:
            raise _lib.BrsError("dropout > 0 is not supported by the B200 NCF engines (the stock configs use 0.0)")
        self.loss = torch.nn.BCELoss()  # attribute kept for API parity (gmf.py:57, ncf.py:92)
        self.batch_size = int(config["model"]["batch_size"])
        super(_NcfEngine, self).__init__(config)
        self.model.to(self.device)

    def _tables(self):
        raise NotImplementedError

    def _bind(self):
        m, dev, opt = self.model, self.device, self.optimizer
        utabs, itabs, n_users, n_items = self._tables()
        b = self.batch_size
        self._user = EntityState(n_users, utabs, opt, b, dev)
        self._item = EntityState(n_items, itabs, opt, b, dev)
        self._dense = []  # (name, weight, grad, state)
        sd = dict(m.named_parameters())
        names = []
        if self.KIND != _lib.NCF_GMF:
            names += [f"fc_layers.{3 * l + 1}" for l in range(m.n_layers)]
        self._fc_names = list(names)
        for base in names + ["affine_output"]:
            for suffix in (".weight", ".bias"):
                w = sd[base + suffix].data
                self._dense.append((base + suffix, w, torch.zeros_like(w), opt.add_param(base + suffix, w)))
        self._ws = torch.zeros(_lib.STEP_WS_BYTES, dtype=torch.uint8, device=dev)
        self._out = torch.zeros(4, dtype=torch.float32, device=dev)
        self._max_batch = 0
        self._alloc_workspace(b)
        m._engine = self

    def _alloc_workspace(self, max_batch):
        """Activation / gradient buffers sized for the largest batch seen so far."""
        if max_batch <= self._max_batch:
            return
        m, dev = self.model, self.device
        self._max_batch = int(max_batch)
        grew = self._user.ensure_capacity(max_batch)
        grew = self._item.ensure_capacity(max_batch) or grew
        emb = m.emb_dim
        c = _lib.NcfModel()
        c.kind, c.emb_dim = self.KIND, emb
        c.user, c.item = self._user.struct, self._item.struct
        self._bufs = []

        def buf(cols):
            t = torch.zeros((self._max_batch, cols), dtype=torch.float32, device=dev)
            self._bufs.append(t)
            return _lib.ptr(t)

        dense = {n: dense_param(w, g, st) for n, w, g, st in self._dense}
        if self.KIND == _lib.NCF_GMF:
            c.n_layers, c.mlp_dim = 0, 0
        else:
            n_layers = m.n_layers
            c.n_layers = n_layers
            c.mlp_dim = emb * (2 ** (n_layers - 1))
            width = 2 * c.mlp_dim
            for l in range(n_layers + 1):
                c.act[l] = buf(width >> l)
                c.dact[l] = buf(width >> l)
            self._wt = []
            for l, base in enumerate(self._fc_names):
                c.fc_weight[l] = dense[base + ".weight"]
                c.fc_bias[l] = dense[base + ".bias"]
                wt = torch.zeros(((width >> l), (width >> l) // 2), dtype=torch.float32, device=dev)
                self._wt.append(wt)  # W^T scratch for the tensor-core dgrad
                c.fc_weight_t[l] = _lib.ptr(wt)
            if self.KIND == _lib.NCF_NEUMF:
                c.mfv = buf(emb)
        c.out_weight = dense["affine_output.weight"]
        c.out_bias = dense["affine_output.bias"]
        c.dz = buf(1)
        c.max_batch = self._max_batch
        c.ws = _lib.ptr(self._ws)
        self._cmodel = c

    @staticmethod
    def _raise_status(status):
        if status == 0:
            return
        if int(status) & 1:
            raise IndexError("index out of range in self")
        raise _lib.BrsError("touched-row list overflow (internal capacity error)")

    def scores(self, users, items):
        lib = _lib.load()
        users, items = as_index(users, self.device), as_index(items, self.device)
        if users.numel() != items.numel():
            raise ValueError("users and items must have the same length")
        out = torch.full((users.numel(),), float("nan"), dtype=torch.float32, device=self.device)
        _lib.check(lib.brs_ncf_predict(self._cmodel, _lib.ptr(users), _lib.ptr(items), users.numel(), _lib.ptr(out),
                                       self._stream()), "brs_ncf_predict")
        self._check_predict()
        return out

    def train_single_batch(self, users, items, ratings):
        """ncf.py:100-120 / gmf.py:60-80 / mlp.py:75-98: (users, items, ratings) -> loss float."""
        assert hasattr(self, "model"), "Please specify the exact model !"
        lib = _lib.load()
        users, items = as_index(users, self.device), as_index(items, self.device)
        ratings = as_float(ratings, self.device)
        b = users.numel()
        if items.numel() != b or ratings.numel() != b:
            raise ValueError("users / items / ratings must have the same length")
        self._alloc_workspace(b)
        _lib.check(lib.brs_ncf_fwd_bwd(self._cmodel, _lib.ptr(users), _lib.ptr(items), _lib.ptr(ratings), b,
                                       self._stream()), "brs_ncf_fwd_bwd")
        _lib.check(lib.brs_ncf_apply(self._cmodel, self.optimizer.desc, b, _lib.ptr(self._out), self._stream()),
                   "brs_ncf_apply")
        loss, _, status = _lib.step_record(self._out)
        self._raise_status(status)
        return loss

    def train_batches(self, users, items, ratings):
        lib = _lib.load()
        users, items = as_index(users, self.device), as_index(items, self.device)
        ratings = as_float(ratings, self.device)
        n, b = users.numel(), self.batch_size
        if n == 0:
            return np.zeros((0, 4), dtype=np.float32)
        self._alloc_workspace(min(n, b))
        n_batches = (n + b - 1) // b
        out = torch.zeros((n_batches, 4), dtype=torch.float32, device=self.device)
        _lib.check(lib.brs_ncf_train_batches(self._cmodel, self.optimizer.desc, _lib.ptr(users), _lib.ptr(items),
                                             _lib.ptr(ratings), n, b, _lib.ptr(out), self._stream()),
                   "brs_ncf_train_batches")
        res = out.cpu().numpy()
        self._raise_status(_lib.step_records_status(res))
        return res

    def train_an_epoch(self, train_loader, epoch_id):
        """ncf.py:122-140: loader yields (user, item, rating); ratings are cast to float."""
        assert hasattr(self, "model"), "Please specify the exact model !"
        t0 = time.time()
        self.model.train()
        total_loss, loss = 0, 0.0
        fast = self._epoch_tensors(train_loader)
        if fast is not None:
            for l in self.train_batches(*fast)[:, 0].tolist():
                loss = l
                total_loss += l
        else:
            for batch_id, batch in enumerate(train_loader):
                user, item, rating = batch[0], batch[1], batch[2]
                loss = self.train_single_batch(user, item, rating.float())
                total_loss += loss
        print("[Training Epoch {}], Loss {}".format(epoch_id, loss if self.KIND == _lib.NCF_NEUMF else total_loss))
        self.writer.add_scalar("model/loss", total_loss, epoch_id)
        print("Execute [train_an_epoch] method costing %.2f ms" % ((time.time() - t0) * 1000))

    def _epoch_tensors(self, loader):
        ds = getattr(loader, "dataset", None)
        names = ("user_tensor", "item_tensor", "target_tensor")  # RatingDataset, data_loaders.py:4-27
        if ds is None or not all(hasattr(ds, n) for n in names):
            return None
        if getattr(loader, "drop_last", False) or getattr(loader, "batch_size", None) != self.batch_size:
            return None
        batches = loader_index_batches(loader)
        if batches is None:
            return None
        order = torch.as_tensor([i for b in batches for i in b], dtype=torch.int64).to(self.device)
        u, i, r = (getattr(ds, n).to(self.device) for n in names)
        return u.index_select(0, order), i.index_select(0, order), r.float().index_select(0, order)


class GMFEngine(_NcfEngine):
    """Drop-in for beta_rec.models.gmf.GMFEngine (gmf.py:51-100)."""

    KIND = _lib.NCF_GMF

    def __init__(self, config):
        self.model = GMF(config["model"])
        self._common_init(config)
        self._bind()

    def _tables(self):
        m = self.model
        return ([("embedding_user.weight", m.embedding_user.weight.data)],
                [("embedding_item.weight", m.embedding_item.weight.data)], m.num_users, m.num_items)


class MLPEngine(_NcfEngine):
    """Drop-in for beta_rec.models.mlp.MLPEngine (mlp.py:66-116)."""

    KIND = _lib.NCF_MLP

    def __init__(self, config):
        self.model = MLP(config["model"])
        self._common_init(config)
        self._bind()

    def _tables(self):
        m = self.model
        return ([("embedding_user.weight", m.embedding_user.weight.data)],
                [("embedding_item.weight", m.embedding_item.weight.data)], m.n_users, m.n_items)


class NeuMFEngine(_NcfEngine):
    """Drop-in for beta_rec.models.ncf.NeuMFEngine (ncf.py:82-193)."""

    KIND = _lib.NCF_NEUMF

    def __init__(self, config):
        self.config = config
        self.model = NeuMF(config["model"])
        self._common_init(config)
        if self.config["model"]["model"] == "ncf_pre" if "model" in self.config["model"] else False:
            self.load_pretrain_weights()
        else:
            self.init_weights()
        self._bind()

    def _tables(self):
        m = self.model
        return ([("embedding_user_mlp.weight", m.embedding_user_mlp.weight.data),
                 ("embedding_user_mf.weight", m.embedding_user_mf.weight.data)],
                [("embedding_item_mlp.weight", m.embedding_item_mlp.weight.data),
                 ("embedding_item_mf.weight", m.embedding_item_mf.weight.data)], m.n_users, m.n_items)

    def init_weights(self):
        """ncf.py:142-154 (the item MLP table is never re-initialised there either)."""
        nn.init.normal_(self.model.embedding_user_mf.weight, std=0.01)
        nn.init.normal_(self.model.embedding_item_mf.weight, std=0.01)
        nn.init.normal_(self.model.embedding_user_mlp.weight, std=0.01)
        nn.init.normal_(self.model.embedding_user_mlp.weight, std=0.01)
        for m1 in self.model.fc_layers:
            if isinstance(m1, nn.Linear):
                nn.init.xavier_uniform_(m1.weight)
        nn.init.kaiming_uniform_(self.model.affine_output.weight, a=1, nonlinearity="sigmoid")

    def load_pretrain_weights(self):
        """ncf.py:156-193: initialise from trained GMF + MLP checkpoints (copied into
        the existing storage so kernel pointers stay valid)."""
        cfg = self.config
        gmf_model = GMF(cfg["model"])
        gmf_sd = torch.load(os.path.join(cfg["system"]["model_save_dir"], cfg["model"]["gmf_config"]["save_name"]),
                            map_location=self.device)
        gmf_model.load_state_dict(gmf_sd)
        mlp_model = MLP(cfg["model"])
        mlp_sd = torch.load(os.path.join(cfg["system"]["model_save_dir"], cfg["model"]["mlp_config"]["save_name"]),
                            map_location=self.device)
        mlp_model.load_state_dict(mlp_sd)
        gmf_model.to(self.device)
        mlp_model.to(self.device)
        with torch.no_grad():
            self.model.embedding_user_mf.weight.copy_(gmf_model.embedding_user.weight)
            self.model.embedding_item_mf.weight.copy_(gmf_model.embedding_item.weight)
            self.model.embedding_user_mlp.weight.copy_(mlp_model.embedding_user.weight)
            self.model.embedding_item_mlp.weight.copy_(mlp_model.embedding_item.weight)
            for m1, m2 in zip(self.model.fc_layers, mlp_model.fc_layers):
                if isinstance(m1, nn.Linear) and isinstance(m2, nn.Linear):
                    m1.weight.copy_(m2.weight)
                    m1.bias.copy_(m2.bias)
            self.model.affine_output.weight.copy_(
                0.5 * torch.cat([mlp_model.affine_output.weight, gmf_model.affine_output.weight], dim=-1))
            self.model.affine_output.bias.copy_(0.5 * (mlp_model.affine_output.bias + gmf_model.affine_output.bias))
