"""Normalised bipartite adjacency built on the GPU: the mirror of ``BaseData.create_adj_mat``
(beta_rec/data/base_data.py:337-360) + ``normalized_adj_single`` (beta_rec/utils/common_util.py:24-41) +
``sparse_mx_to_torch_sparse_tensor`` (beta_rec/recommenders/lightgcn.py:15-23).

``build_norm_adj(users, items, n_users, n_items)`` returns a :class:`GpuAdjacency` holding the CSR of
``D^-1 (A + I)`` (and of its transpose, with the forward-edge map the backward SpMM needs) in device memory;
pass it as ``config["model"]["norm_adj"]`` to ``LightGCNEngine`` in place of the reference's torch sparse
tensor (which is still accepted).  ``to_torch_sparse()`` gives the reference's own representation back.
"""
import ctypes

import numpy as np
import torch

from . import _lib


class GpuAdjacency(object):
    """CSR arrays (int32 / fp32, on ``device``) of an [n, n] row-normalised adjacency with a symmetric
    pattern: row_ptr / col / val, and for the transpose (same row_ptr / col) val_t and edge_id_t."""

    def __init__(self, n_users, n_items, row_ptr, col, val, val_t, edge_id_t):
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.n = self.n_users + self.n_items
        self.row_ptr, self.col, self.val, self.val_t, self.edge_id_t = row_ptr, col, val, val_t, edge_id_t
        self.nnz = int(col.numel())
        self.device = col.device

    def csr_tensors(self):
        """The dict LightGCNEngine binds (same keys as engines.lightgcn.coo_to_csr)."""
        return {"row_ptr": self.row_ptr, "col": self.col, "val": self.val, "row_ptr_t": self.row_ptr, "col_t": self.col,
                "val_t": self.val_t, "edge_id_t": self.edge_id_t}

    def to_torch_sparse(self):
        """Coalesced torch sparse COO tensor, the reference's representation (recommenders/lightgcn.py:15-23)."""
        counts = (self.row_ptr[1:] - self.row_ptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(self.n, device=self.device), counts)
        idx = torch.stack([rows, self.col.long()])
        return torch.sparse_coo_tensor(idx, self.val, (self.n, self.n)).coalesce()


def _ids(x, dev):
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x)).astype(np.int64, copy=False))
    return x.to(device=dev, dtype=torch.int64).contiguous().view(-1)


def build_norm_adj(users, items, n_users, n_items, mean=False, device="cuda"):
    """norm_adj = D^-1 (A + I) (``mean=False``) or mean_adj = D^-1 A (``mean=True``) of the bipartite graph of
    the training interactions (u, i); duplicate interactions collapse (R[u, i] = 1)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.BrsError("beta_recsys_b200.graph needs a CUDA device (there is no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    u, i = _ids(users, dev), _ids(items, dev)
    if u.numel() != i.numel():
        raise ValueError("users / items must have the same length")
    e, n = u.numel(), int(n_users) + int(n_items)
    nbytes = lib.brs_adj_workspace_bytes(e, int(n_users), int(n_items))
    if nbytes <= 0:
        raise _lib.BrsError("adjacency too large for int32 CSR")
    bound = 2 * e + n
    with torch.cuda.device(dev):
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        row_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        col = torch.empty(bound, dtype=torch.int32, device=dev)
        val = torch.empty(bound, dtype=torch.float32, device=dev)
        val_t = torch.empty(bound, dtype=torch.float32, device=dev)
        eid = torch.empty(bound, dtype=torch.int32, device=dev)
        nnz = torch.zeros(1, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.brs_adj_build(_lib.ptr(u), _lib.ptr(i), e, int(n_users), int(n_items), 0 if mean else 1, _lib.ptr(ws),
                                     nbytes, _lib.ptr(row_ptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(val_t), _lib.ptr(eid),
                                     _lib.ptr(nnz), st), "brs_adj_build")
        status = ctypes.c_uint32(0)
        _lib.check(lib.brs_adj_status(_lib.ptr(ws), e, int(n_users), int(n_items), ctypes.byref(status), st), "brs_adj_status")
        if status.value:
            raise IndexError("an interaction lies outside [0, n_users) x [0, n_items)")
        k = int(nnz.item())
    # shrink to nnz (the bound counts duplicate interactions twice)
    return GpuAdjacency(n_users, n_items, row_ptr, col[:k].clone(), val[:k].clone(), val_t[:k].clone(), eid[:k].clone())
