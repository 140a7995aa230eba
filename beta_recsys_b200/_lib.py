"""ctypes binding of libbrs_b200.so (the C ABI declared in include/brs_b200.h).

There is NO fallback: if the library is missing or a call fails, the product
raises.  PyTorch is used by the callers only to own device memory and streams;
no torch type crosses this boundary -- only raw device pointers and sizes.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbrs_b200.so")

BRS_OK = 0
SGD, ADAM, RMSPROP = 0, 1, 2
DENSE, TOUCHED_ROWS = 0, 1
MAX_ENTITY_TABLES = 4
STEP_WS_BYTES = 256
OPT_KINDS = {"sgd": SGD, "adam": ADAM, "rmsprop": RMSPROP}


class BrsError(RuntimeError):
    pass


class Opt(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mode", C.c_int32), ("lr", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("eps", C.c_double), ("alpha", C.c_double)]


class Table(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("grad", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("n_rows", C.c_int64), ("dim", C.c_int32), ("pad_", C.c_int32)]


class Rowset(C.Structure):
    _fields_ = [("slot_map", C.c_void_p), ("list", C.c_void_p), ("count", C.c_void_p), ("n_rows", C.c_int64),
                ("capacity", C.c_int32), ("pad_", C.c_int32)]


class Entity(C.Structure):
    _fields_ = [("rows", Rowset), ("n_tables", C.c_int32), ("pad_", C.c_int32), ("table", Table * MAX_ENTITY_TABLES)]


class DenseParam(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("grad", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("numel", C.c_int64)]


class MfPlan(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("bytes", C.c_int64), ("batch_capacity", C.c_int64),
                ("user_capacity", C.c_int32), ("item_capacity", C.c_int32)]


class MfModel(C.Structure):
    _fields_ = [("user", Entity), ("item", Entity), ("global_bias", DenseParam), ("ws", C.c_void_p),
                ("user_rows_alt", Rowset), ("item_rows_alt", Rowset), ("plan", MfPlan * 2), ("user_stage", C.c_void_p)]


NCF_GMF, NCF_MLP, NCF_NEUMF = 0, 1, 2
NCF_MAX_LAYERS = 6


class NcfModel(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_layers", C.c_int32), ("emb_dim", C.c_int32), ("mlp_dim", C.c_int32),
                ("user", Entity), ("item", Entity),
                ("fc_weight", DenseParam * NCF_MAX_LAYERS), ("fc_bias", DenseParam * NCF_MAX_LAYERS),
                ("fc_weight_t", C.c_void_p * NCF_MAX_LAYERS), ("out_weight", DenseParam), ("out_bias", DenseParam),
                ("act", C.c_void_p * (NCF_MAX_LAYERS + 1)), ("dact", C.c_void_p * (NCF_MAX_LAYERS + 1)),
                ("mfv", C.c_void_p), ("dz", C.c_void_p), ("max_batch", C.c_int64), ("ws", C.c_void_p)]


LGCN_MAX_LAYERS = 6


class Csr(C.Structure):
    _fields_ = [("row_ptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p), ("edge_id", C.c_void_p),
                ("n_rows", C.c_int64), ("nnz", C.c_int64)]


class LightGCNModel(C.Structure):
    _fields_ = [("n_users", C.c_int64), ("n_items", C.c_int64), ("dim", C.c_int32), ("n_layers", C.c_int32),
                ("decay", C.c_float), ("pad_", C.c_float), ("adj", Csr), ("adj_t", Csr),
                ("emb", C.c_void_p * (LGCN_MAX_LAYERS + 1)), ("d", C.c_void_p), ("g", C.c_void_p * 2),
                ("param", DenseParam), ("ws", C.c_void_p)]


MAX_RANKS = 8
IPC_HANDLE_BYTES = 64


class MfPeerTables(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "user_emb", "item_emb", "user_bias", "item_bias", "g_user_emb", "g_item_emb", "g_user_bias", "g_item_bias",
        "user_bits", "item_bits")]


class PeerSync(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("flags", C.c_void_p * MAX_RANKS),
                ("partials", C.c_void_p * MAX_RANKS)]


class MfSharded(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("n_users", C.c_int64), ("n_items", C.c_int64),
                ("local_users", C.c_int64), ("local_items", C.c_int64), ("stage", MfModel), ("peers", C.c_void_p),
                ("own", MfPeerTables), ("pull_user_emb", C.c_void_p), ("pull_item_emb", C.c_void_p),
                ("pull_user_bias", C.c_void_p), ("pull_item_bias", C.c_void_p)]


EXTRA_STRUCTS = {"brs_mf_plan": MfPlan, "brs_csr": Csr, "brs_lightgcn_model": LightGCNModel, "brs_ncf_model": NcfModel, "brs_mf_peer_tables": MfPeerTables, "brs_peer_sync": PeerSync,
                 "brs_mf_sharded": MfSharded}

_P = C.c_void_p
_PROTOTYPES = {
    # name: (restype, argtypes)
    "brs_abi_version": (C.c_int, []),
    "brs_strerror": (C.c_char_p, [C.c_int]),
    "brs_last_cuda_error": (C.c_char_p, []),
    "brs_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int64)]),
    "brs_mf_bpr_fwd_bwd": (C.c_int, [C.POINTER(MfModel), _P, _P, _P, C.c_int64, C.c_float, _P]),
    "brs_mf_bpr_prepare": (C.c_int, [C.POINTER(MfModel), _P, _P, _P, C.c_int64, _P]),
    "brs_mf_bpr_fwd_bwd_prepared": (C.c_int, [C.POINTER(MfModel), _P, _P, _P, C.c_int64, C.c_float, _P]),
    "brs_debug_set_l2_policy": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "brs_mf_bce_fwd_bwd": (C.c_int, [C.POINTER(MfModel), _P, _P, _P, C.c_int64, C.c_float, _P]),
    "brs_mf_apply": (C.c_int, [C.POINTER(MfModel), C.POINTER(Opt), C.c_int64, _P, _P]),
    "brs_mf_train_batches": (C.c_int, [C.POINTER(MfModel), C.POINTER(Opt), C.c_int32, _P, _P, _P, C.c_int64,
                                       C.c_int64, C.c_float, _P, _P]),
    "brs_mf_train_batches_host": (C.c_int, [C.POINTER(MfModel), C.POINTER(Opt), C.c_int32, _P, _P, _P, C.c_int64,
                                       C.c_int64, C.c_float, _P, _P]),
    "brs_mf_predict": (C.c_int, [C.POINTER(MfModel), _P, _P, C.c_int64, _P, _P]),
    "brs_mf_plan_bytes": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "brs_mf_plan_build": (C.c_int, [C.POINTER(MfModel), C.c_int32, C.c_int32, _P, _P, _P, C.c_int64, _P]),
    "brs_mf_step_planned": (C.c_int, [C.POINTER(MfModel), C.c_int32, C.POINTER(Opt), C.c_int32, C.c_int64, C.c_float,
                                      _P, _P]),
    "brs_debug_set_mf_rows_only": (C.c_int, [C.c_int]),
    "brs_mf_step": (C.c_int, [C.POINTER(MfModel), C.POINTER(Opt), C.c_int32, _P, _P, _P, C.c_int64, C.c_float, _P,
                              _P]),
    "brs_ncf_fwd_bwd": (C.c_int, [C.POINTER(NcfModel), _P, _P, _P, C.c_int64, _P]),
    "brs_ncf_apply": (C.c_int, [C.POINTER(NcfModel), C.POINTER(Opt), C.c_int64, _P, _P]),
    "brs_ncf_predict": (C.c_int, [C.POINTER(NcfModel), _P, _P, C.c_int64, _P, _P]),
    "brs_ncf_train_batches": (C.c_int, [C.POINTER(NcfModel), C.POINTER(Opt), _P, _P, _P, C.c_int64, C.c_int64, _P,
                                        _P]),
    "brs_mlp_fwd_tc": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P]),
    "brs_set_gemm_backend": (C.c_int, [C.c_int]),
    "brs_mlp_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P]),
    "brs_mlp_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "brs_spmm_csr": (C.c_int, [C.POINTER(Csr), _P, C.c_float, _P, _P, C.c_int32, _P]),
    "brs_lightgcn_propagate": (C.c_int, [C.POINTER(LightGCNModel), _P, C.c_float, _P]),
    "brs_lightgcn_fwd_bwd": (C.c_int, [C.POINTER(LightGCNModel), _P, C.c_float, _P, _P, _P, C.c_int64, _P]),
    "brs_lightgcn_tail": (C.c_int, [C.POINTER(LightGCNModel), _P, _P, _P, C.c_int64, C.c_int64, _P]),
    "brs_lightgcn_reg_grad": (C.c_int, [C.POINTER(LightGCNModel), _P, _P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, _P]),
    "brs_lightgcn_apply": (C.c_int, [C.POINTER(LightGCNModel), C.POINTER(Opt), C.c_int64, _P, _P]),
    "brs_lightgcn_scores": (C.c_int, [C.POINTER(LightGCNModel), _P, _P, C.c_int64, _P, _P]),
    "brs_rows_assign": (C.c_int, [C.POINTER(Rowset), _P, C.c_int64, _P, _P]),
    "brs_rows_scatter_grad": (C.c_int, [C.POINTER(Entity), C.c_int32, _P, C.c_int64, _P, C.c_float, _P]),
    "brs_rows_read_grad": (C.c_int, [C.POINTER(Entity), C.c_int32, _P, C.c_int64, _P, _P]),
    "brs_rows_sgd": (C.c_int, [C.POINTER(Entity), C.c_int32, C.c_double, _P]),
    "brs_rows_adam": (C.c_int, [C.POINTER(Entity), C.c_int32, C.POINTER(Opt), C.c_int64, _P]),
    "brs_dense_adam_sweep": (C.c_int, [C.POINTER(Entity), C.c_int32, C.POINTER(Opt), C.c_int64, _P]),
    "brs_dense_params_step": (C.c_int, [C.POINTER(DenseParam), C.c_int32, C.POINTER(Opt), C.c_int64, _P]),
    "brs_shm_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "brs_shm_free": (C.c_int, [_P]),
    "brs_ipc_get_handle": (C.c_int, [_P, C.POINTER(C.c_uint8)]),
    "brs_ipc_open_handle": (C.c_int, [C.POINTER(C.c_uint8), C.POINTER(C.c_void_p)]),
    "brs_ipc_close_handle": (C.c_int, [_P]),
    "brs_peer_barrier": (C.c_int, [C.POINTER(PeerSync), C.c_uint64, _P, _P]),
    "brs_mf_sharded_bpr_fwd_bwd": (C.c_int, [C.POINTER(MfSharded), _P, _P, _P, C.c_int64, C.c_int64, C.c_float, _P]),
    "brs_mf_sharded_apply": (C.c_int, [C.POINTER(MfSharded), C.POINTER(Opt), C.c_int64, _P, _P]),
    "brs_debug_set_shard_mode": (C.c_int, [C.c_int]),
    "brs_mf_sharded_push": (C.c_int, [C.POINTER(MfSharded), C.POINTER(Opt), _P]),
    "brs_mf_sharded_step": (C.c_int, [C.POINTER(MfSharded), C.POINTER(PeerSync), C.POINTER(Opt), _P, _P, _P, C.c_int64,
                                      C.c_int64, C.c_float, C.c_uint64, _P, _P]),
    "brs_mf_sharded_train_batches": (C.c_int, [C.POINTER(MfSharded), C.POINTER(PeerSync), C.POINTER(Opt), _P, _P, _P,
                                               C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_uint64, _P, _P]),
    "brs_mf_sharded_train_batches_host": (C.c_int, [C.POINTER(MfSharded), C.POINTER(PeerSync), C.POINTER(Opt), _P, _P, _P,
                                               C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_uint64, _P, _P]),
    "brs_route_triples": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, _P, _P, _P, _P, _P]),
    "brs_gather": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int64, _P, _P]),
    "brs_scatter_add": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int64, _P, C.c_float, _P]),
    "brs_gather_sgd_update": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int64, C.c_float, _P, _P]),
    "brs_adj_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64, C.c_int64]),
    "brs_adj_build": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _P, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "brs_adj_status": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_uint32), _P]),
    "brs_pairset_bytes": (C.c_int64, [C.c_int64]),
    "brs_pairset_build": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64, _P]),
    "brs_sample_negatives": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32, C.c_uint64, _P, _P]),
    "brs_pairset_status": (C.c_int, [_P, C.POINTER(C.c_uint32), _P]),
    "brs_rank_metrics_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64, C.c_int64]),
    "brs_rank_metrics": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, _P, C.c_int64, _P, _P]),
}

_lib = None


def load():
    """Load (once) and return the shared library; raise loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BrsError(
            "libbrs_b200.so not found at %s -- build it with `python -m beta_recsys_b200.build` "
            "(there is no CPU or PyTorch fallback)" % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if lib.brs_abi_version() != 2:
        raise BrsError("libbrs_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status, what=""):
    if status != BRS_OK:
        lib = load()
        msg = lib.brs_strerror(status).decode()
        if status == -3:
            msg += ": " + lib.brs_last_cuda_error().decode()
        raise BrsError("%s failed: %s" % (what or "libbrs_b200 call", msg))


def step_record(out):
    """brs_step_out -> (loss, regularizer, status) from a float32[4] device tensor (one D2H copy, the
    reference's two .item() syncs) -- status is an int32 bit mask stored in the third word."""
    rec = out.detach().cpu().numpy()
    return float(rec[0]), float(rec[1]), int(rec.view("int32")[2])


def step_records_status(res):
    """OR of the status words of a float32 [n, 4] numpy array of brs_step_out records."""
    import numpy as np

    if res.shape[0] == 0:
        return 0
    return int(np.bitwise_or.reduce(np.ascontiguousarray(res).view(np.int32)[:, 2]))


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def make_opt(optimizer, lr, mode=DENSE):
    if optimizer not in OPT_KINDS:
        raise ValueError("unsupported optimizer %r (sgd | adam | rmsprop)" % (optimizer,))
    return Opt(OPT_KINDS[optimizer], mode, float(lr), 0.9, 0.999, 1e-8, 0.99)
