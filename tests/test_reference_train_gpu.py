"""BASELINE.json configs[0] (plumbing) on the GPU: the reference's OWN ``MatrixFactorization(config).train(data)``
-- recommender, TrainEngine._train loop, early stopping, checkpointing, EvalEngine worker threads, BaseData and
its instance_bpr_loader -- runs unmodified, with ``beta_recsys_b200.install()`` having rebound the engine class
and ``evaluate``.  Needs the reference package (``/root/reference`` in the build container, ``baseline/_ref`` on
the GPU box: oracle/install_ref.sh) and the import shim for its six missing third-party modules."""
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.needs_reference]

# configs/mf_default.json of the reference (its default hyper-parameters), with max_epoch cut to 3
MF_DEFAULT = {
    "system": {"root_dir": "../", "log_dir": "logs/", "result_dir": "results/", "process_dir": "processes/",
               "checkpoint_dir": "checkpoints/", "dataset_dir": "datasets/", "run_dir": "runs/", "tune_dir": "tune_results/",
               "device": "gpu", "seed": 2020, "metrics": ["ndcg", "precision", "recall", "map"], "k": [5, 10, 20],
               "valid_metric": "ndcg", "valid_k": 10, "result_file": "mf_result.csv", "save_mode": "average"},
    "dataset": {"dataset": "ml_100k", "data_split": "leave_one_out", "download": False, "random": False, "test_rate": 0.2,
                "by_user": False, "n_test": 10, "n_negative": 100,
                "result_col": ["dataset", "data_split", "test_rate", "n_negative"]},
    "model": {"model": "MF", "config_id": "default", "emb_dim": 64, "num_negative": 4, "batch_size": 400, "batch_eval": True,
              "dropout": 0.0, "optimizer": "adam", "loss": "bpr", "lr": 0.05, "reg": 0.001, "max_epoch": 3, "max_n_update": 20,
              "save_name": "mf.model",
              "result_col": ["model", "emb_dim", "batch_size", "dropout", "optimizer", "loss", "lr", "reg"]},
    "tunable": [{"name": "loss", "type": "choice", "values": ["bce", "bpr"]}],
}


def synthetic_split(n_users=943, n_items=1682, n_inter=30000, n_neg=50, seed=2020):
    """ML-100k-shaped interactions (cfg 1), leave-one-out: the last interaction of a user is the test
    positive, the one before the validation positive; both evaluated against n_neg sampled non-interacted items."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    pu = np.arange(1, n_users + 1, dtype=np.float64) ** -1.05
    pi = np.arange(1, n_items + 1, dtype=np.float64) ** -1.05
    u = rng.permutation(n_users)[rng.choice(n_users, n_inter, p=pu / pu.sum())]
    i = rng.permutation(n_items)[rng.choice(n_items, n_inter, p=pi / pi.sum())]
    df = pd.DataFrame({"col_user": u, "col_item": i}).drop_duplicates().reset_index(drop=True)
    df["col_rating"] = 1.0
    df["col_timestamp"] = np.arange(len(df))
    df = df.sort_values(["col_user", "col_timestamp"])
    rank_from_end = df.groupby("col_user").cumcount(ascending=False)
    n_per_user = df.groupby("col_user")["col_item"].transform("count")
    test_pos = df[(rank_from_end == 0) & (n_per_user >= 3)]
    valid_pos = df[(rank_from_end == 1) & (n_per_user >= 3)]
    train = df.drop(test_pos.index).drop(valid_pos.index).reset_index(drop=True)
    seen = df.groupby("col_user")["col_item"].apply(set).to_dict()

    def with_negatives(pos):
        rows = []
        for uu, ii in zip(pos["col_user"], pos["col_item"]):
            rows.append((uu, ii, 1.0))
            cand = rng.choice(n_items, 3 * n_neg)
            cand = [c for c in dict.fromkeys(cand.tolist()) if c not in seen[uu]][:n_neg]
            rows.extend((uu, c, 0.0) for c in cand)
        return pd.DataFrame(rows, columns=["col_user", "col_item", "col_rating"]).astype(
            {"col_user": np.int64, "col_item": np.int64, "col_rating": np.float64})

    keep = ["col_user", "col_item", "col_rating"]
    return train[keep].copy(), with_negatives(valid_pos), with_negatives(test_pos)


def run_reference_train(tmp_path, device, use_install):
    from oracle import ref_shim

    ref_shim.install()
    import beta_recsys_b200
    from beta_rec.data.base_data import BaseData
    from beta_rec.recommenders.matrix_factorization import MatrixFactorization

    cfg = json.loads(json.dumps(MF_DEFAULT))
    cfg["system"]["root_dir"] = str(tmp_path) + "/"
    # TrainEngine.get_device (core/train_engine.py:37-57): a one-character value is a gpu id -> "cuda:<id>"
    # (its "cuda:#" branch calls int(":0") and cannot be used); "cpu" stays "cpu"
    cfg["system"]["device"] = "cpu" if device == "cpu" else device.replace("cuda:", "")
    cfg_file = tmp_path / "mf_default.json"
    cfg_file.write_text(json.dumps(cfg))
    train, valid, test = synthetic_split()
    patched = beta_recsys_b200.install(strict=True) if use_install else []
    try:
        data = BaseData((train, [valid], [test]))
        rec = MatrixFactorization({"config_file": str(cfg_file), "device": device, "root_dir": str(tmp_path) + "/"})
        result = rec.train(data)
        deadline = time.time() + 120
        while rec.eval_engine.n_worker > 0 and time.time() < deadline:  # the metric workers are threads (eval_engine.py:504-515)
            time.sleep(0.2)
        return rec, result, patched
    finally:
        if use_install:
            beta_recsys_b200.uninstall()


def test_reference_matrix_factorization_train_runs_on_the_b200_engines(tmp_path):
    import torch

    from beta_recsys_b200 import engines

    buf = io.StringIO()
    with redirect_stdout(buf):
        rec, result, patched = run_reference_train(tmp_path, "cuda:0", use_install=True)
    sys.stdout, sys.stderr = sys.__stdout__, sys.__stderr__  # the reference installs its own Logger objects there
    import glob

    log = buf.getvalue() + "".join(open(f, errors="ignore").read() for f in glob.glob(str(tmp_path / "logs" / "*")))
    assert ("beta_rec.recommenders.matrix_factorization", "MFEngine") in patched
    assert ("beta_rec.core.eval_engine", "evaluate") in patched
    assert isinstance(rec.engine, engines.MFEngine) and rec.engine.device.type == "cuda"
    # three epochs trained by our engine (its own log line), each followed by the reference's evaluation
    assert log.count("[Training Epoch") >= 3 and "Execute [train_an_epoch] method costing" in log
    assert rec.eval_engine.n_worker == 0
    # the GPU `evaluate` fed the reference's early-stopping bookkeeping: a real NDCG@10 came back
    assert 0.0 < result["valid_metric"] <= 1.0, result
    # the reference's checkpoint path holds a state_dict in the reference layout
    assert os.path.exists(result["model_save_dir"]), result
    sd = torch.load(result["model_save_dir"], map_location="cpu")
    assert set(sd) == {"global_bias", "user_emb.weight", "item_emb.weight", "user_bias.weight", "item_bias.weight"}
    assert tuple(sd["user_emb.weight"].shape) == (rec.config["model"]["n_users"], 64)
    # learning happened: a trained model ranks the validation positives above chance (1 of 51 candidates: NDCG@10 ~ 0.09)
    assert result["valid_metric"] > 0.10, result
