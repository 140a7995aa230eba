"""CPU restatement (numpy, per-user loops) of the reference's ranking metrics -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
path (beta_recsys_b200/eval.py -> csrc/eval_kernels.cu) never does.

Follows beta_rec/utils/evaluation.py:
  merge_ranking_true_pred  :459-534   true rows with rating >= 1; users common to both frames; per user the
                                      k highest predictions (nlargest keeps the FIRST of equal scores, then
                                      rank(method="first", ascending=False)); hits = top-k rows whose
                                      (user, item) is a true row; actual = true rows per user
  get_top_k_items          :755-785
  precision_at_k           :537-583   sum_u hit_u / k / n_users
  recall_at_k              :586-629   sum_u hit_u / actual_u / n_users
  ndcg_at_k                :632-690   dcg_u = sum_hits 1/ln(1 + rank); idcg_u = sum_{r<=min(actual_u,k)} 1/ln(1 + r)
  map_at_k                 :693-752   sum_hits (cumcount + 1) / rank / actual_u
and beta_rec/core/eval_engine.py:49-87 (evaluate: the prediction frame is the data frame's own rows).

Pinned on the known answers of the reference's own tests (/root/reference/tests/test_evaluation.py:
ndcg 0.38172, map 0.23613, precision 0.26666, recall 0.37777 at k = 10; 1.0 / 0.6 for perfect predictions;
0.0 without hits) in tests/test_eval_oracle.py.  The reference functions themselves no longer run in this
container (pandas 3.0 drops the grouping column in groupby.apply).
"""
import numpy as np

METRICS = ("ndcg", "map", "precision", "recall")


def rank_metrics(true_users, true_items, true_ratings, pred_users, pred_items, pred_scores, k=10):
    """All four ranking metrics at k.  Returns a dict metric -> float (0.0 when there is no hit at all)."""
    true_users = np.asarray(true_users)
    true_items = np.asarray(true_items)
    pred_users = np.asarray(pred_users)
    pred_items = np.asarray(pred_items)
    pred_scores = np.asarray(pred_scores, dtype=np.float64)
    keep = np.asarray(true_ratings, dtype=np.float64) >= 1  # evaluation.py:492
    true_users, true_items = true_users[keep], true_items[keep]
    truth = {}
    for u, i in zip(true_users.tolist(), true_items.tolist()):
        truth.setdefault(u, []).append(i)
    rows = {}
    for idx, u in enumerate(pred_users.tolist()):
        rows.setdefault(u, []).append(idx)
    common = [u for u in truth if u in rows]  # evaluation.py:495-497
    n_users = len(common)
    sums = dict.fromkeys(METRICS, 0.0)
    n_hits = 0
    for u in common:
        idx = np.asarray(rows[u])
        # k largest, ties in original row order (nlargest keep="first"), then rank 1..k (method="first")
        order = idx[np.argsort(-pred_scores[idx], kind="stable")][:k]
        true_set = set(truth[u])
        actual = len(truth[u])
        hit_ranks = [r + 1 for r, j in enumerate(order) if pred_items[j] in true_set]
        if not hit_ranks:
            continue  # users without a hit are absent from df_hit_count: they add 0 to every sum
        n_hits += len(hit_ranks)
        sums["precision"] += len(hit_ranks) / k
        sums["recall"] += len(hit_ranks) / actual
        dcg = sum(1.0 / np.log1p(r) for r in hit_ranks)
        idcg = sum(1.0 / np.log1p(r) for r in range(1, min(actual, k) + 1))
        sums["ndcg"] += dcg / idcg
        sums["map"] += sum((c + 1) / r for c, r in enumerate(hit_ranks)) / actual
    if n_hits == 0 or n_users == 0:
        return dict.fromkeys(METRICS, 0.0)
    return {m: sums[m] / n_users for m in METRICS}


def evaluate(users, items, ratings, predictions, metrics, k_li):
    """core/eval_engine.py:49-87 for the ranking metrics: the prediction rows are the data rows."""
    if not isinstance(k_li, list):
        k_li = [k_li]
    out = {}
    for k in k_li:
        r = rank_metrics(users, items, ratings, users, items, predictions, k)
        for m in metrics:
            out["%s@%d" % (m, k)] = r[m]
    return out
