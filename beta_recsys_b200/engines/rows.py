"""Device-side scratch for one id space (users / items): gradient accumulators,
touched-row bitmap + list, optimizer state -- allocated ONCE by PyTorch and
handed to the kernels as raw pointers (include/brs_b200.h: brs_entity)."""
import torch

from .. import _lib


class EntityState(object):
    def __init__(self, n_rows, tables, optimizer, capacity, device):
        """tables: list of (name, weight tensor [n_rows, dim]) sharing this id space."""
        assert 1 <= len(tables) <= _lib.MAX_ENTITY_TABLES
        self.n_rows = int(n_rows)
        self.names = [n for n, _ in tables]
        self.weights = [w for _, w in tables]
        self.capacity = int(max(1, min(self.n_rows, capacity)))
        self.slot_map = torch.full((self.n_rows,), -1, dtype=torch.int32, device=device)
        self.count = torch.zeros(1, dtype=torch.int32, device=device)
        self.states = []
        for name, w in tables:
            assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.shape[0] == self.n_rows
            self.states.append(optimizer.add_param(name, w))
        # alternate bookkeeping (slot map / list / counter) so that the epoch loop can claim slots for
        # batch b+1 while batch b is being applied; the gradient scratch itself is shared
        self.slot_map_alt = torch.full((self.n_rows,), -1, dtype=torch.int32, device=device)
        self.count_alt = torch.zeros(1, dtype=torch.int32, device=device)
        self._alloc_scratch()
        self.struct = self._make_struct()

    def alt_rowset(self):
        return _lib.Rowset(_lib.ptr(self.slot_map_alt), _lib.ptr(self.list_alt), _lib.ptr(self.count_alt), self.n_rows,
                           self.capacity, 0)

    def _alloc_scratch(self):
        """Compact gradient scratch [capacity, dim] per table + the slot list: the only
        per-step gradient storage (the reference materialises table-sized dense grads)."""
        dev = self.slot_map.device
        self.list = torch.zeros(self.capacity, dtype=torch.int32, device=dev)
        self.list_alt = torch.zeros(self.capacity, dtype=torch.int32, device=dev)
        self.grads = [torch.zeros((self.capacity, w.shape[1]), dtype=torch.float32, device=dev) for w in self.weights]

    def _make_struct(self):
        e = _lib.Entity()
        e.rows = _lib.Rowset(_lib.ptr(self.slot_map), _lib.ptr(self.list), _lib.ptr(self.count), self.n_rows,
                             self.capacity, 0)
        e.n_tables = len(self.weights)
        for k, (w, g, st) in enumerate(zip(self.weights, self.grads, self.states)):
            e.table[k] = _lib.Table(_lib.ptr(w), _lib.ptr(g), _lib.ptr(st.get("m")), _lib.ptr(st.get("v")),
                                    self.n_rows, int(w.shape[1]), 0)
        return e

    def ensure_capacity(self, capacity):
        """Grow the touched list if a larger batch arrives (rows would be lost otherwise)."""
        capacity = int(max(1, min(self.n_rows, capacity)))
        if capacity > self.capacity:
            self.capacity = capacity
            self._alloc_scratch()  # between steps the scratch is all-zero, so nothing to carry over
            self.struct = self._make_struct()
            return True
        return False


def dense_param(weight, grad, state):
    return _lib.DenseParam(_lib.ptr(weight), _lib.ptr(grad), _lib.ptr(state.get("m")), _lib.ptr(state.get("v")),
                           weight.numel())


def as_index(t, device):
    """int64, contiguous, on `device` -- the reference's LongTensor batches (data_loaders.py:43-49)."""
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if t.dtype != torch.int64:
        t = t.long()
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous().view(-1)


def as_float(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if t.dtype != torch.float32:
        t = t.float()
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous().view(-1)


def loader_index_batches(loader):
    """The index batches a torch DataLoader WOULD yield, without materialising
    samples one __getitem__ at a time (the reference spends 58-75% of its step
    time there).  Consumes the global RNG exactly like
    ``iter(loader)`` does, so shuffles are bit-identical to the reference's.
    Returns None when the loader is not a plain single-process DataLoader."""
    try:
        from torch.utils.data import DataLoader
    except Exception:  # pragma: no cover
        return None
    if not isinstance(loader, DataLoader) or loader.num_workers != 0 or loader.batch_sampler is None:
        return None
    if loader.collate_fn is not torch.utils.data.dataloader.default_collate:
        return None
    it = iter(loader)
    sampler_iter = getattr(it, "_sampler_iter", None)
    if sampler_iter is None:
        return None
    return [list(b) for b in sampler_iter]
