"""Ranking evaluation on the GPU: the mirror of beta_rec/core/eval_engine.py:49-87 (``evaluate``) and
beta_rec/utils/evaluation.py:459-752 (``ndcg_at_k`` / ``map_at_k`` / ``precision_at_k`` / ``recall_at_k``).

``evaluate(data_df, predictions, metrics, k_li)`` has the reference's signature and result keys
("ndcg@10", ...); ``install()`` rebinds ``beta_rec.core.eval_engine.evaluate`` to it, so the reference's
``train_eval_worker`` / ``test_eval_worker`` run unmodified on top of csrc/eval_kernels.cu.  The frame's
columns go to the device as tensors; nothing is ranked on the host and there is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib

USER_COL, ITEM_COL, RATING_COL, PREDICTION_COL = "col_user", "col_item", "col_rating", "col_prediction"
RANKING_METRICS = ("ndcg", "map", "precision", "recall")


def _dev(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise _lib.BrsError("beta_recsys_b200.eval needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _ids(x, dev):
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x)).astype(np.int64, copy=False))
    return x.to(device=dev, dtype=torch.int64).contiguous()


def _vals(x, dev):
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    return x.to(device=dev, dtype=torch.float32).contiguous()


def rank_metrics(true_users, true_items, true_ratings, pred_users, pred_items, pred_scores, k=10, n_user_ids=None,
                 device=None):
    """ndcg / map / precision / recall at k of the prediction rows against the true rows (rating >= 1),
    exactly as evaluation.py:459-752 defines them (ties keep the frame's row order).  Returns a dict."""
    dev = _dev(device)
    lib = _lib.load()
    tu, ti, tr = _ids(true_users, dev), _ids(true_items, dev), _vals(true_ratings, dev)
    pu, pi, ps = _ids(pred_users, dev), _ids(pred_items, dev), _vals(pred_scores, dev)
    if not (tu.numel() == ti.numel() == tr.numel()) or not (pu.numel() == pi.numel() == ps.numel()):
        raise ValueError("user / item / value columns must have the same length")
    if n_user_ids is None:
        hi = -1
        if tu.numel():
            hi = max(hi, int(tu.max().item()))
        if pu.numel():
            hi = max(hi, int(pu.max().item()))
        n_user_ids = hi + 1
    n_user_ids = max(int(n_user_ids), 1)
    with torch.cuda.device(dev):
        nbytes = lib.brs_rank_metrics_workspace_bytes(tu.numel(), pu.numel(), n_user_ids)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty(8, dtype=torch.float64, device=dev)
        _lib.check(lib.brs_rank_metrics(_lib.ptr(tu), _lib.ptr(ti), _lib.ptr(tr), tu.numel(), _lib.ptr(pu), _lib.ptr(pi),
                                        _lib.ptr(ps), pu.numel(), n_user_ids, int(k), _lib.ptr(ws), nbytes, _lib.ptr(out),
                                        torch.cuda.current_stream(dev).cuda_stream), "brs_rank_metrics")
        res = out.cpu().numpy()
    if int(res[6]) != 0:
        raise IndexError("user id outside [0, %d) or item id outside [0, 2^32) in the evaluation frame" % n_user_ids)
    n_users, n_hits = res[4], res[5]
    if n_users == 0 or n_hits == 0:  # evaluation.py: `if df_hit.shape[0] == 0: return 0.0`
        return dict.fromkeys(RANKING_METRICS, 0.0)
    return {"ndcg": float(res[0] / n_users), "map": float(res[1] / n_users), "precision": float(res[2] / n_users),
            "recall": float(res[3] / n_users)}


def evaluate(data_df, predictions, metrics, k_li):
    """core/eval_engine.py:49-87: ``data_df`` holds col_user / col_item / col_rating, ``predictions`` one
    score per row.  Returns {"<metric>@<k>": value}.  Ranking metrics only (rmse / mae / rsquared are
    rating-prediction metrics outside this path)."""
    bad = [m for m in metrics if m not in RANKING_METRICS]
    if bad:
        raise NotImplementedError("metrics %s are not ranking metrics; supported: %s" % (bad, list(RANKING_METRICS)))
    users = data_df[USER_COL].to_numpy()
    items = data_df[ITEM_COL].to_numpy()
    ratings = data_df[RATING_COL].to_numpy()
    if not isinstance(k_li, list):
        k_li = [k_li]
    dev = _dev()
    u, i, r = _ids(users, dev), _ids(items, dev), _vals(ratings, dev)
    p = _vals(predictions, dev)
    out = {}
    for k in k_li:
        res = rank_metrics(u, i, r, u, i, p, k=k, device=dev)
        for m in metrics:
            out["%s@%s" % (m, k)] = res[m]
    return out
