"""GPU negative sampler (csrc/sample_kernels.cu) bit-exact against oracle/sample_oracle.py, and the loaders
it feeds through MFEngine.train_an_epoch."""
import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from oracle import sample_oracle as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("num_negative", [1, 4])
def test_sampler_bit_exact_vs_oracle(num_negative):
    from beta_recsys_b200 import sampling as G

    rng = np.random.default_rng(num_negative)
    n_users, n_items = 300, 120
    tu, ti = rng.integers(0, n_users, 5000), rng.integers(0, n_items, 5000)
    s = G.InteractionSet(tu, ti, n_users, n_items)
    got = s.sample_negatives(tu, num_negative, seed=2020).cpu().numpy()
    want = S.sample_negatives(tu, ti, tu, n_items, num_negative, seed=2020)
    assert np.array_equal(got, want)
    u, p, n = G.bpr_triples(tu, ti, n_users, n_items, seed=9)
    ou, op, on = S.bpr_triples(tu, ti, n_items, seed=9)
    assert np.array_equal(n.cpu().numpy(), on) and np.array_equal(p.cpu().numpy(), op)
    u, i, r = G.bce_samples(tu, ti, np.ones(len(tu), np.float32), n_users, n_items, 3, seed=5)
    ou, oi, orr = S.bce_samples(tu, ti, np.ones(len(tu), np.float32), n_items, 3, seed=5)
    assert np.array_equal(u.cpu().numpy(), ou) and np.array_equal(i.cpu().numpy(), oi) and np.array_equal(r.cpu().numpy(), orr)


def test_sampler_errors():
    from beta_recsys_b200 import sampling as G

    with pytest.raises(IndexError):
        G.InteractionSet([0, 5], [1, 1], 5, 3)
    s = G.InteractionSet(np.zeros(4, dtype=np.int64), np.arange(4), 2, 4)
    with pytest.raises(ValueError):
        s.sample_negatives([0], 1, seed=0)


def test_sampler_large_and_properties():
    """1M interactions over 200k x 50k: no negative is a positive of its user (checked on the device)."""
    from beta_recsys_b200 import sampling as G

    g = torch.Generator().manual_seed(1)
    n_users, n_items, n = 200_000, 50_000, 1_000_000
    tu = torch.randint(0, n_users, (n,), generator=g)
    ti = torch.randint(0, n_items, (n,), generator=g)
    u, p, neg = G.bpr_triples(tu, ti, n_users, n_items, seed=3)
    assert int(neg.min()) >= 0 and int(neg.max()) < n_items
    key_pos = torch.unique(u * n_items + p)
    key_neg = u * n_items + neg
    assert not bool(torch.isin(key_neg, key_pos).any())
    # roughly uniform over items
    cnt = torch.bincount(neg, minlength=n_items).float()
    assert abs(float(cnt.mean()) - n / n_items) < 1e-3 and float(cnt.max()) < 60


def test_bpr_loader_feeds_train_an_epoch_like_the_reference_loader():
    """The loader built on the device drives MFEngine.train_an_epoch's fast path; same shuffle as a CPU
    DataLoader over the same tensors (global torch RNG), checked against the oracle."""
    from beta_recsys_b200 import sampling as G
    from beta_recsys_b200.engines import MFEngine

    rng = np.random.default_rng(5)
    n_users, n_items, bsz = 90, 70, 64
    tu, ti = rng.integers(0, n_users, 700), rng.integers(0, n_items, 700)
    loader = G.instance_bpr_loader((tu, ti), bsz, "cuda:0", n_users, n_items, seed=4)
    cfg = {"model": dict(device_str="cuda:0", n_users=n_users, n_items=n_items, emb_dim=16, batch_size=bsz, optimizer="sgd",
                         lr=0.05, loss="bpr"), "system": {"run_dir": "/tmp/brs_test"}}
    torch.manual_seed(0)
    eng = MFEngine(cfg)
    p = {k: v.detach().cpu().numpy().copy() for k, v in eng.model.state_dict().items()}
    st = O.new_opt_state(p, "sgd")
    ds = loader.dataset
    cols = [c.cpu().numpy() for c in (ds.user_tensor, ds.pos_item_tensor, ds.neg_item_tensor)]
    torch.manual_seed(123)
    order = [list(b) for b in iter(torch.utils.data.DataLoader(ds, batch_size=bsz, shuffle=True))._sampler_iter]
    torch.manual_seed(123)
    eng.train_an_epoch(loader, epoch_id=0)
    for b in order:
        O.mf_train_single_batch(p, st, tuple(c[b] for c in cols), "bpr", "sgd", 0.05, 0.0)
    for k, v in eng.model.state_dict().items():
        # the biases start at 0 and sum opposite-sign terms: absolute budget for them (cf. test_mf_gpu.Scale)
        scale = max(np.abs(p[k]).max(), 1e-2)
        err = np.abs(v.cpu().numpy() - p[k]).max() / scale
        assert err <= 1e-5, (k, err)
