"""TEST INFRASTRUCTURE ONLY -- torch-CPU port of the reference training step.

``cf_oracle.py`` is the closed-form checker; this file restates the SAME
reference step with the SAME PyTorch ops the reference executes (dense
``nn.Embedding`` gradients through autograd, ``torch.optim`` over every row),
so that timing it on the host cores measures what the reference's own CPU
path costs.  ``/root/reference`` cannot travel to the GPU box, so this port is
what ``bench.py`` times as ``cpu_baseline`` (kind "port") and as
``--impl reference``.  It is validated against the live reference in the build
container by ``tests/test_oracle_golden.py::test_torch_port_matches_reference``
(bit-identical: same ops, same order).

Reference call sites restated: beta_rec/models/mf.py:32-55,92-119;
beta_rec/models/torch_engine.py:23-39,92-121; beta_rec/models/gmf.py:29-36;
beta_rec/models/ncf.py:52-71,100-120; beta_rec/models/lightgcn.py:27-78,119-191.
"""
import torch
import torch.nn.functional as F


def make_optimizer(params, optimizer, lr):
    """ModelEngine.set_optimizer (beta_rec/models/torch_engine.py:23-39)."""
    if optimizer == "sgd":
        return torch.optim.SGD(params, lr=lr)
    if optimizer == "adam":
        return torch.optim.Adam(params, lr=lr)
    if optimizer == "rmsprop":
        return torch.optim.RMSprop(params, lr=lr)
    raise ValueError(optimizer)


class MFPort(object):
    """MF + MFEngine (beta_rec/models/mf.py)."""

    KEYS = ("global_bias", "user_emb.weight", "item_emb.weight", "user_bias.weight", "item_bias.weight")

    def __init__(self, state, optimizer="sgd", lr=0.05, loss="bpr", reg=0.0):
        self.p = {k: torch.nn.Parameter(torch.as_tensor(state[k]).clone().float()) for k in self.KEYS}
        self.opt = make_optimizer([self.p[k] for k in self.KEYS], optimizer, lr)
        self.loss, self.reg = loss, reg

    def forward(self, users, items):
        p = self.p
        u = F.embedding(users, p["user_emb.weight"])
        ub = F.embedding(users, p["user_bias.weight"])
        i = F.embedding(items, p["item_emb.weight"])
        ib = F.embedding(items, p["item_bias.weight"])
        scores = torch.sigmoid(
            torch.sum(torch.mul(u, i).squeeze(), dim=1) + ub.squeeze() + ib.squeeze() + p["global_bias"]
        )
        reg = ((u ** 2).sum() + (i ** 2).sum() + (ub ** 2).sum() + (ib ** 2).sum()) / u.size()[0]
        return scores, reg

    def train_single_batch(self, batch):
        self.opt.zero_grad()
        if self.loss == "bpr":
            users, pos, neg = batch
            ps, pr = self.forward(users, pos)
            ns, nr = self.forward(users, neg)
            loss = -torch.mean(F.logsigmoid(ps - ns))
            reg = pr + nr
        elif self.loss == "bce":
            users, items, ratings = batch
            s, reg = self.forward(users, items)
            loss = torch.nn.BCELoss()(s, ratings)
        else:
            raise RuntimeError(f"Unsupported loss type {self.loss}, try other options: 'bpr' or 'bce'")
        (loss + self.reg * reg).backward()
        self.opt.step()
        return loss.item(), reg.item()

    def state(self):
        return {k: v.detach().numpy().copy() for k, v in self.p.items()}


def _fc_keys(n_layers):
    return [f"fc_layers.{3 * l + 1}" for l in range(n_layers)]


class NeuMFPort(object):
    """NeuMF + NeuMFEngine (beta_rec/models/ncf.py), dropout 0."""

    def __init__(self, state, n_layers, optimizer="adam", lr=1e-3):
        self.n_layers = n_layers
        self.p = {k: torch.nn.Parameter(torch.as_tensor(v).clone().float()) for k, v in state.items()}
        self.opt = make_optimizer(list(self.p.values()), optimizer, lr)

    def forward(self, users, items):
        p = self.p
        um = F.embedding(users, p["embedding_user_mlp.weight"])
        im = F.embedding(items, p["embedding_item_mlp.weight"])
        uf = F.embedding(users, p["embedding_user_mf.weight"])
        if_ = F.embedding(items, p["embedding_item_mf.weight"])
        x = torch.relu(torch.cat([um, im], dim=-1))  # ReLU after the leading Dropout (ncf.py:64-66)
        for k in _fc_keys(self.n_layers):
            x = torch.relu(F.linear(x, p[k + ".weight"], p[k + ".bias"]))
        vec = torch.cat([x, uf * if_], dim=-1)
        return torch.sigmoid(F.linear(vec, p["affine_output.weight"], p["affine_output.bias"]))

    def train_single_batch(self, users, items, ratings):
        self.opt.zero_grad()
        loss = torch.nn.BCELoss()(self.forward(users, items).view(-1), ratings)
        loss.backward()
        self.opt.step()
        return loss.item()

    def state(self):
        return {k: v.detach().numpy().copy() for k, v in self.p.items()}


class GMFPort(object):
    """GMF + GMFEngine (beta_rec/models/gmf.py)."""

    def __init__(self, state, optimizer="adam", lr=1e-3):
        self.p = {k: torch.nn.Parameter(torch.as_tensor(v).clone().float()) for k, v in state.items()}
        self.opt = make_optimizer(list(self.p.values()), optimizer, lr)

    def forward(self, users, items):
        p = self.p
        u = F.embedding(users, p["embedding_user.weight"])
        i = F.embedding(items, p["embedding_item.weight"])
        return torch.sigmoid(F.linear(u * i, p["affine_output.weight"], p["affine_output.bias"]))

    def train_single_batch(self, users, items, ratings):
        self.opt.zero_grad()
        loss = torch.nn.BCELoss()(self.forward(users, items).view(-1), ratings)
        loss.backward()
        self.opt.step()
        return loss.item()

    def state(self):
        return {k: v.detach().numpy().copy() for k, v in self.p.items()}


class LightGCNPort(object):
    """LightGCN + LightGCNEngine (beta_rec/models/lightgcn.py).  ``adj`` is the
    torch sparse COO Ahat; the edge-dropout keep mask is passed in so that the
    checker and the CUDA path consume the same mask the reference would draw."""

    def __init__(self, state, adj, n_layers, decay, keep_prob, optimizer="adam", lr=0.05):
        self.p = {k: torch.nn.Parameter(torch.as_tensor(v).clone().float()) for k, v in state.items()}
        self.opt = make_optimizer(list(self.p.values()), optimizer, lr)
        self.adj = adj.coalesce()
        self.n_layers, self.decay, self.keep_prob = n_layers, decay, keep_prob

    def dropout(self, keep_mask):
        x = self.adj
        index = x.indices().t()[keep_mask]
        values = x.values()[keep_mask] / self.keep_prob
        return torch.sparse_coo_tensor(index.t(), values, x.size())

    def train_single_batch(self, batch, keep_mask=None):
        p = self.p
        nu = p["user_embedding.weight"].shape[0]
        self.opt.zero_grad()
        adj = self.adj if keep_mask is None else self.dropout(keep_mask)
        e = torch.cat((p["user_embedding.weight"], p["item_embedding.weight"]), dim=0)
        embs = [e]
        for _ in range(self.n_layers):
            e = torch.sparse.mm(adj, e)
            embs.append(e)
        ebar = torch.mean(torch.stack(embs, dim=1), dim=1)
        ue, ie = ebar[:nu], ebar[nu:]
        users, pos, neg = batch
        u, pi, nj = ue[users], ie[pos], ie[neg]
        ps = torch.sum(u * pi, dim=1)
        ns = torch.sum(u * nj, dim=1)
        u0 = F.embedding(users, p["user_embedding.weight"])
        p0 = F.embedding(pos, p["item_embedding.weight"])
        n0 = F.embedding(neg, p["item_embedding.weight"])
        reg = 0.5 * (u0.norm(2).pow(2) + p0.norm(2).pow(2) + n0.norm(2).pow(2)) / float(len(users))
        loss = torch.mean(F.softplus(ns - ps)) + reg * self.decay
        loss.backward()
        self.opt.step()
        return loss.item()

    def state(self):
        return {k: v.detach().numpy().copy() for k, v in self.p.items()}
