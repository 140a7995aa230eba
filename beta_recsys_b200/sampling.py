"""Negative sampling and training loaders on the GPU: the mirror of ``BaseData.instance_bpr_loader`` /
``instance_bce_loader`` (beta_rec/data/base_data.py:182-253) and of the two dataset wrappers in
beta_rec/data/data_loaders.py:4-53.

The reference materialises, per user, the Python set of every item the user has NOT interacted with
(O(U * I)) and draws with ``random.sample``; it cannot build BASELINE configs 2-5.  Here the training
interactions go into a device hash set once (csrc/sample_kernels.cu) and every row draws its negatives by
rejection from a counter-based stream -- the same distribution (uniform over non-interacted items, a row's
negatives pairwise distinct), reproducible from ``seed`` and independent of the launch shape.  The loaders
returned are plain ``torch.utils.data.DataLoader(shuffle=True)`` objects over datasets with the reference's
attribute names, so ``MFEngine.train_an_epoch`` takes its device fast path and the shuffle consumes the
global torch RNG exactly like the reference's loader does.
"""
import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from . import _lib

USER_COL, ITEM_COL, RATING_COL = "col_user", "col_item", "col_rating"


class RatingDataset(Dataset):
    """data_loaders.py:4-27."""

    def __init__(self, user_tensor, item_tensor, target_tensor):
        self.user_tensor = user_tensor
        self.item_tensor = item_tensor
        self.target_tensor = target_tensor

    def __getitem__(self, index):
        return self.user_tensor[index], self.item_tensor[index], self.target_tensor[index]

    def __len__(self):
        return self.user_tensor.size(0)


class PairwiseNegativeDataset(Dataset):
    """data_loaders.py:30-53."""

    def __init__(self, user_tensor, pos_item_tensor, neg_item_tensor):
        self.user_tensor = user_tensor
        self.pos_item_tensor = pos_item_tensor
        self.neg_item_tensor = neg_item_tensor

    def __getitem__(self, index):
        return self.user_tensor[index], self.pos_item_tensor[index], self.neg_item_tensor[index]

    def __len__(self):
        return self.user_tensor.size(0)


def _dev(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.BrsError("beta_recsys_b200.sampling needs a CUDA device (there is no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _ids(x, dev):
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x)).astype(np.int64, copy=False))
    return x.to(device=dev, dtype=torch.int64).contiguous().view(-1)


class InteractionSet(object):
    """The (user, item) pairs of the training interactions as a device hash set."""

    def __init__(self, users, items, n_users, n_items, device="cuda"):
        self.device = _dev(device)
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.users, self.items = _ids(users, self.device), _ids(items, self.device)
        if self.users.numel() != self.items.numel():
            raise ValueError("users / items must have the same length")
        self.n_pairs = self.users.numel()
        lib = _lib.load()
        with torch.cuda.device(self.device):
            self._buf = torch.empty(lib.brs_pairset_bytes(self.n_pairs), dtype=torch.uint8, device=self.device)
            _lib.check(lib.brs_pairset_build(_lib.ptr(self.users), _lib.ptr(self.items), self.n_pairs, self.n_users,
                                             self.n_items, _lib.ptr(self._buf), self._buf.numel(), self._stream()),
                       "brs_pairset_build")
        self._check()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _check(self):
        import ctypes

        st = ctypes.c_uint32(0)
        _lib.check(_lib.load().brs_pairset_status(_lib.ptr(self._buf), ctypes.byref(st), self._stream()), "brs_pairset_status")
        if st.value & 1:
            raise IndexError("an interaction lies outside [0, n_users) x [0, n_items)")
        if st.value & 2:  # random.sample on an empty population (base_data.py:241)
            raise ValueError("Sample larger than population: a user has interacted with every item")

    def sample_negatives(self, users, num_negative=1, seed=0):
        """int64 [len(users), num_negative] on the device: for row r, ``num_negative`` distinct items that
        ``users[r]`` has not interacted with."""
        users = _ids(users, self.device)
        out = torch.empty((users.numel(), int(num_negative)), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().brs_sample_negatives(_lib.ptr(self._buf), self.n_pairs, _lib.ptr(users), users.numel(),
                                                        self.n_items, int(num_negative), int(seed) & (2 ** 64 - 1),
                                                        _lib.ptr(out), self._stream()), "brs_sample_negatives")
        self._check()
        return out


def bpr_triples(users, items, n_users, n_items, seed=0, device="cuda"):
    """(users, pos_items, neg_items) device LongTensors: instance_bpr_loader's three columns."""
    s = InteractionSet(users, items, n_users, n_items, device)
    return s.users, s.items, s.sample_negatives(s.users, 1, seed).view(-1)


def bce_samples(users, items, ratings, n_users, n_items, num_negative, seed=0, device="cuda"):
    """(users, items, ratings) device tensors: every training row followed by its ``num_negative`` negatives
    with rating 0, the row order of instance_bce_loader (base_data.py:203-210)."""
    s = InteractionSet(users, items, n_users, n_items, device)
    neg = s.sample_negatives(s.users, num_negative, seed)
    if not torch.is_tensor(ratings):
        ratings = torch.from_numpy(np.ascontiguousarray(np.asarray(ratings, dtype=np.float32)))
    r = ratings.to(device=s.device, dtype=torch.float32).view(-1, 1)
    u = s.users.view(-1, 1).expand(-1, num_negative + 1).reshape(-1)
    i = torch.cat([s.items.view(-1, 1), neg], dim=1).reshape(-1)
    t = torch.cat([r, torch.zeros((r.shape[0], num_negative), dtype=torch.float32, device=s.device)], dim=1).reshape(-1)
    return u.contiguous(), i.contiguous(), t.contiguous()


def _columns(train, with_rating=False):
    if hasattr(train, "columns"):  # the reference's DataFrame (base_data.py:220-246)
        cols = [train[USER_COL].to_numpy(), train[ITEM_COL].to_numpy()]
        if with_rating:
            cols.append(train[RATING_COL].to_numpy())
        return cols
    return list(train)


def instance_bpr_loader(train, batch_size, device, n_users, n_items, seed=0):
    """BaseData.instance_bpr_loader (base_data.py:218-253) for a DataFrame (col_user / col_item) or a
    (users, items) pair of arrays."""
    u, i = _columns(train)[:2]
    users, pos, neg = bpr_triples(u, i, n_users, n_items, seed, device)
    dataset = PairwiseNegativeDataset(users, pos, neg)
    print(f"Making PairwiseNegativeDataset of length {len(dataset)}")
    return DataLoader(dataset, batch_size=batch_size, shuffle=True)


def instance_bce_loader(train, batch_size, device, num_negative, n_users, n_items, seed=0):
    """BaseData.instance_bce_loader (base_data.py:182-216)."""
    u, i, r = _columns(train, with_rating=True)[:3]
    users, items, ratings = bce_samples(u, i, r, n_users, n_items, num_negative, seed, device)
    dataset = RatingDataset(users, items, ratings)
    print(f"Making RatingDataset of length {len(dataset)}")
    return DataLoader(dataset, batch_size=batch_size, shuffle=True)
