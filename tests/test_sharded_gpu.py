"""GPU tests of the row-sharded multi-GPU MF path (peer-memory kernels, flag barrier,
triple routing).  world_size 1 runs on any GPU box; world_size 2/4/8 need that many
GPUs (skipped otherwise).  Parity target: the sharded engines, each fed its own batch,
equal the oracle run on the concatenated global batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cf_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _state(rng, nu, ni, d):
    return {
        "global_bias": np.array([0.03], dtype=np.float32),
        "user_emb.weight": rng.normal(0, 0.1, (nu, d)).astype(np.float32),
        "item_emb.weight": rng.normal(0, 0.1, (ni, d)).astype(np.float32),
        "user_bias.weight": rng.normal(0, 0.1, (nu, 1)).astype(np.float32),
        "item_bias.weight": rng.normal(0, 0.1, (ni, 1)).astype(np.float32),
    }


def _zipf(rng, n, size, a=1.05):
    p = np.arange(1, n + 1, dtype=np.float64) ** (-a)
    p /= p.sum()
    return rng.permutation(n)[rng.choice(n, size=size, p=p)].astype(np.int64)


def _worker(rank, world, port, route, optimizer, q, mode=0, feed="step"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from beta_recsys_b200 import _lib
        from beta_recsys_b200.sharded import ShardedMFEngine

        _lib.check(_lib.load().brs_debug_set_shard_mode(mode))  # 0: by world size, 1: per-sample peer gathers, 2: pull + staging
        nu, ni, d, bsz, lr, steps = 5003, 1999, 128, 1024, 0.05, (6 if feed == "host" else 3)  # 6 > ring depth
        rng = np.random.default_rng(7)  # same on every rank: the global model and all batches
        p = _state(rng, nu, ni, d)
        batches = [[(_zipf(rng, nu, bsz), _zipf(rng, ni, bsz), rng.integers(0, ni, bsz)) for _ in range(world)]
                   for _ in range(steps)]
        cfg = {"model": dict(device_str="cuda:%d" % rank, n_users=nu, n_items=ni, emb_dim=d, batch_size=bsz,
                             optimizer=optimizer, lr=lr, loss="bpr")}
        eng = ShardedMFEngine(cfg, route=route, state=p)
        st = O.new_opt_state(p, optimizer)
        if feed == "host":  # the whole epoch from pinned HOST arrays through the streaming C loop
            mine = [torch.from_numpy(np.concatenate([batches[t][rank][c] for t in range(steps)])).pin_memory()
                    for c in range(3)]
            rec = eng.train_batches(*mine)
            assert rec.shape == (steps, 4) and not rec[:, 2].any()
        for t in range(steps):
            if feed == "host":
                loss, reg = float(rec[t, 0]), float(rec[t, 1])
            else:
                loss, reg = eng.train_single_batch(tuple(torch.from_numpy(x).cuda() for x in batches[t][rank]))
            gu = np.concatenate([b[0] for b in batches[t]])
            gp = np.concatenate([b[1] for b in batches[t]])
            gn = np.concatenate([b[2] for b in batches[t]])
            if optimizer == "sgd" or t == 0:
                ol, orr = O.mf_train_single_batch(p, st, (gu, gp, gn), "bpr", optimizer, lr, 0.0)
                assert abs(loss - ol) <= 1e-5 * max(1, abs(ol)), (t, loss, ol)
                assert abs(reg - orr) <= 1e-5 * max(1, abs(orr)), (t, reg, orr)
        got = eng.gather_state()
        if optimizer == "sgd":
            for k in p:
                scale = max(np.abs(p[k]).max(), 0.05)
                err = np.abs(got[k].astype(np.float64) - p[k]).max() / scale
                assert err <= 1e-5, (k, err)
        # replicas of the replicated parameter are bit-identical across ranks
        gb = [torch.empty(1, device="cuda") for _ in range(world)]
        dist.all_gather(gb, eng.global_bias)
        assert all(torch.equal(gb[0], x) for x in gb)
        # scratch is clean again on every rank
        for name in ("user_slot", "item_slot"):
            assert bool((eng._stage[name] == -1).all().item())
        for name in ("g_user_emb", "g_item_emb", "g_user_bias", "g_item_bias"):
            assert float(eng.arena.tensor(name).abs().max().item()) == 0.0
        for name in ("user_bits", "item_bits"):
            assert int(eng.arena.tensor(name).abs().max().item()) == 0
        assert float(eng._stage["s_user_emb"].abs().max().item()) == 0.0
        eng.close()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


def _run(world, route, optimizer, mode=0, feed="step"):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = _free_port()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")  # children import tests.*
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, route, optimizer, q, mode, feed)) for r in range(world)]
    # children re-import this module by name from the parent's sys.path: keep the repo root first (an
    # earlier test may have put the reference checkout, which has its own `tests` package, in front)
    saved = sys.path[:]
    sys.path[:] = [root] + [x for x in saved if x != root]
    try:
        for p in procs:
            p.start()
    finally:
        sys.path[:] = saved
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


@pytest.mark.timeout(300)
@pytest.mark.parametrize("optimizer", ["sgd", "adam"])
def test_sharded_world1_matches_oracle(optimizer):
    _run(1, "none", optimizer)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("route", ["none", "owner"])
def test_sharded_world2_matches_oracle_on_the_global_batch(route):
    _run(2, route, "sgd")


@pytest.mark.timeout(300)
def test_sharded_world2_adam_first_step():
    _run(2, "none", "adam")


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,optimizer", [(1, "adam"), (2, "sgd"), (2, "adam")])
def test_sharded_staged_mode_matches_oracle(world, optimizer):
    """mode 2: every unique row pulled once into the staging tables, fused kernel on local memory"""
    _run(world, "none", optimizer, mode=2)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [1, 2])
def test_sharded_epoch_from_host_memory(world):
    """brs_mf_sharded_train_batches_host: per-rank index arrays stay in pinned host memory"""
    _run(world, "none", "sgd", feed="host")


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [4, 8])
def test_sharded_world_4_8(world):
    _run(world, "none", "sgd")


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_route_triples_kernel_is_bit_exact_vs_oracle(world):
    from beta_recsys_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(world)
    for n in (1, 255, 256, 257, 40000):
        u, p, ng = rng.integers(0, 10**6, n), rng.integers(0, 10**5, n), rng.integers(0, 10**5, n)
        tu, tp, tn = (torch.from_numpy(x).cuda() for x in (u, p, ng))
        ou, op_, on = torch.empty_like(tu), torch.empty_like(tp), torch.empty_like(tn)
        counts = torch.empty(world, dtype=torch.int64, device="cuda")
        _lib.check(lib.brs_route_triples(tu.data_ptr(), tp.data_ptr(), tn.data_ptr(), n, world, ou.data_ptr(),
                                         op_.data_ptr(), on.data_ptr(), counts.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
        c, ru, rp, rn, _ = O.route_triples(u, p, ng, world)
        assert np.array_equal(counts.cpu().numpy(), c)
        assert np.array_equal(ou.cpu().numpy(), ru) and np.array_equal(op_.cpu().numpy(), rp)
        assert np.array_equal(on.cpu().numpy(), rn)


# --------------------------------------------------------------------------- #
# sharded checkpoint: save (gather -> reference state_dict) / resume (scatter), optimizer state included
# --------------------------------------------------------------------------- #
def _ckpt_worker(rank, world, port, optimizer, q, path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from beta_recsys_b200.sharded import ShardedMFEngine

        nu, ni, d, bsz, lr = 1003, 499, 64, 512, 0.05
        rng = np.random.default_rng(11)
        p = _state(rng, nu, ni, d)
        batches = [[(_zipf(rng, nu, bsz), _zipf(rng, ni, bsz), rng.integers(0, ni, bsz)) for _ in range(world)] for _ in range(3)]
        cfg = {"model": dict(device_str="cuda:%d" % rank, n_users=nu, n_items=ni, emb_dim=d, batch_size=bsz,
                             optimizer=optimizer, lr=lr, loss="bpr", adam_mode="dense")}

        def step(eng, t):
            return eng.train_single_batch(tuple(torch.from_numpy(x).cuda() for x in batches[t][rank]))

        a = ShardedMFEngine(cfg, state=p)
        step(a, 0)
        step(a, 1)
        a.save_checkpoint(path)
        step(a, 2)
        want = a.gather_state()
        want_opt = a.gather_optimizer_state()
        a.close()
        # the file is the reference module's state_dict: keys, shapes, dtypes (models/mf.py:17-30)
        sd = torch.load(path, map_location="cpu")
        assert {k: tuple(v.shape) for k, v in sd.items()} == {
            "global_bias": (1,), "user_emb.weight": (nu, d), "item_emb.weight": (ni, d), "user_bias.weight": (nu, 1),
            "item_bias.weight": (ni, 1)}
        assert all(v.dtype == torch.float32 for v in sd.values())
        b = ShardedMFEngine(cfg)  # fresh random tables
        b.resume_checkpoint(path)
        step(b, 2)
        got, got_opt = b.gather_state(), b.gather_optimizer_state()
        b.close()
        # same step from the same state: only the order of the fp32 gradient atomics may differ -- ~1e-7 relative on
        # g, which lr * m / (sqrt(v) + eps) amplifies without bound where |g| ~ eps (tests/test_oracle_golden.py), so
        # the adaptive optimizers get the documented 2e-3 * lr budget and SGD the tight one
        tol = 1e-6 if optimizer == "sgd" else 2e-3 * lr
        for k in want:
            assert np.abs(got[k] - want[k]).max() <= tol * max(1.0, np.abs(want[k]).max()), k
            assert np.mean(np.abs(got[k] - want[k]) > 1e-6) < 0.01, k  # and only isolated elements use it
        assert set(got_opt) == set(want_opt)
        for idx in want_opt:
            assert float(got_opt[idx]["step"]) == float(want_opt[idx]["step"]) == 3.0
            for name in want_opt[idx]:
                if name != "step":
                    assert torch.allclose(got_opt[idx][name], want_opt[idx][name], rtol=1e-5, atol=1e-9), (idx, name)
        if rank == 0:  # the same file loads into the single-GPU engine
            from beta_recsys_b200.engines import MFEngine

            c1 = {"model": dict(cfg["model"], device_str="cuda:0"), "system": {"run_dir": "/tmp/brs_test"}}
            e1 = MFEngine(c1)
            e1.resume_checkpoint(path)
            for k, v in e1.model.state_dict().items():
                assert torch.equal(v.cpu(), sd[k]), k
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,optimizer", [(1, "adam"), (1, "sgd"), (2, "adam"), (2, "rmsprop")])
def test_sharded_checkpoint_roundtrip(world, optimizer, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = _free_port()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    path = str(tmp_path / "mf_sharded.ckpt")
    procs = [ctx.Process(target=_ckpt_worker, args=(r, world, port, optimizer, q, path)) for r in range(world)]
    saved = sys.path[:]
    sys.path[:] = [root] + [x for x in saved if x != root]
    try:
        for p in procs:
            p.start()
    finally:
        sys.path[:] = saved
    res = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


# --------------------------------------------------------------------------- #
# row-sharded NeuMF (BASELINE configs[2]): every rank its own batch == the oracle on the global batch
# --------------------------------------------------------------------------- #
def _neumf_state(rng, nu, ni, emb, nl):
    w = 2 * emb * 2 ** (nl - 1)
    st = {"embedding_user_mlp.weight": rng.normal(0, 0.1, (nu, w // 2)), "embedding_item_mlp.weight": rng.normal(0, 0.1, (ni, w // 2)),
          "embedding_user_mf.weight": rng.normal(0, 0.1, (nu, emb)), "embedding_item_mf.weight": rng.normal(0, 0.1, (ni, emb)),
          "affine_output.weight": rng.normal(0, 0.3, (1, w // 2 ** nl + emb)), "affine_output.bias": rng.normal(0, 0.1, (1,))}
    for l in range(nl):
        st["fc_layers.%d.weight" % (3 * l + 1)] = rng.normal(0, 0.1, (w >> (l + 1), w >> l))
        st["fc_layers.%d.bias" % (3 * l + 1)] = rng.normal(0, 0.05, (w >> (l + 1),))
    return {k: v.astype(np.float32) for k, v in st.items()}


def _neumf_worker(rank, world, port, optimizer, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from beta_recsys_b200.sharded_ncf import ShardedNeuMFEngine

        nu, ni, emb, nl, bsz, lr = 1003, 499, 16, 2, 256, (0.05 if optimizer == "sgd" else 1e-3)
        rng = np.random.default_rng(5)
        p = _neumf_state(rng, nu, ni, emb, nl)
        steps = 3 if optimizer == "sgd" else 1
        batches = [[(_zipf(rng, nu, bsz), _zipf(rng, ni, bsz), (rng.random(bsz) < 0.3).astype(np.float32)) for _ in range(world)]
                   for _ in range(steps)]
        cfg = {"model": dict(model="ncf_end", device_str="cuda:%d" % rank, n_users=nu, n_items=ni, emb_dim=emb, batch_size=bsz,
                             optimizer=optimizer, lr=lr, dropout=0.0, adam_mode="dense", mlp_config={"n_layers": nl}),
               "system": {"run_dir": "/tmp/brs_test"}}
        import io
        from contextlib import redirect_stdout

        with redirect_stdout(io.StringIO()):
            eng = ShardedNeuMFEngine(cfg, state=p)
        st = O.new_opt_state(p, optimizer)
        for t in range(steps):
            loss = eng.train_single_batch(*[torch.from_numpy(x).cuda() for x in batches[t][rank]])
            g = [np.concatenate([b[c] for b in batches[t]]) for c in range(3)]
            ol = O.neumf_train_single_batch(p, st, g[0], g[1], g[2], nl, optimizer, lr)
            assert abs(loss - ol) <= 1e-5 * max(1.0, abs(ol)), (t, loss, ol)
        got = eng.gather_state()
        assert set(got) == set(p)
        for k in p:
            if optimizer == "sgd":
                scale = max(np.abs(p[k]).max(), 0.05)
                err = np.abs(got[k].astype(np.float64) - p[k]).max() / scale
                assert err <= 2e-5, (k, err)
            else:  # first Adam step: |dw| <= lr whatever the rounding; compare within the documented budget
                assert np.abs(got[k].astype(np.float64) - p[k]).max() <= 2e-3 * lr + 1e-7, k
        with pytest.raises(IndexError):
            bad = batches[0][rank][0].copy()
            bad[3] = nu
            eng.train_single_batch(torch.from_numpy(bad).cuda(), torch.from_numpy(batches[0][rank][1]).cuda(),
                                   torch.from_numpy(batches[0][rank][2]).cuda())
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()[-1800:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,optimizer", [(1, "sgd"), (1, "adam"), (2, "sgd"), (2, "adam"), (4, "sgd")])
def test_sharded_neumf_matches_oracle_on_the_global_batch(world, optimizer):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = _free_port()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_neumf_worker, args=(r, world, port, optimizer, q)) for r in range(world)]
    saved = sys.path[:]
    sys.path[:] = [root] + [x for x in saved if x != root]
    try:
        for p in procs:
            p.start()
    finally:
        sys.path[:] = saved
    res = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


# --------------------------------------------------------------------------- #
# row-partitioned LightGCN (BASELINE configs[3]): every rank its own batch == the oracle on the global batch
# --------------------------------------------------------------------------- #
def _lightgcn_worker(rank, world, port, optimizer, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from beta_recsys_b200.sharded_lightgcn import ShardedLightGCNEngine

        nu, ni, d, L, n_e, bsz, decay, keep = 1501, 733, 64, 3, 30000, 512, 1e-4, 0.6
        lr = 0.05 if optimizer == "sgd" else 0.01
        rng = np.random.default_rng(9)
        pu, pi = np.arange(1, nu + 1) ** -1.0, np.arange(1, ni + 1) ** -1.0
        eu, ei = rng.choice(nu, n_e, p=pu / pu.sum()), rng.choice(ni, n_e, p=pi / pi.sum())
        adj = O.row_normalised_adj(nu, ni, eu, ei)
        coo = adj.tocoo()
        tadj = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([coo.row, coo.col]).astype(np.int64)),
                                       torch.from_numpy(coo.data.astype(np.float32)), (nu + ni, nu + ni))
        p = {"user_embedding.weight": rng.normal(0, 0.1, (nu, d)).astype(np.float32),
             "item_embedding.weight": rng.normal(0, 0.1, (ni, d)).astype(np.float32)}
        cfg = {"model": dict(device_str="cuda:%d" % rank, n_users=nu, n_items=ni, emb_dim=d, layer_size=[d] * L, batch_size=bsz,
                             optimizer=optimizer, lr=lr, regs=[decay], keep_pro=keep, norm_adj=tadj)}
        eng = ShardedLightGCNEngine(cfg, state=p)
        st = O.new_opt_state(p, optimizer)
        steps = 2 if optimizer == "sgd" else 1
        for t in range(steps):
            batches = [(rng.integers(0, nu, bsz), rng.integers(0, ni, bsz), rng.integers(0, ni, bsz)) for _ in range(world)]
            mask = rng.random(adj.nnz) < keep
            loss = eng.train_single_batch(tuple(torch.from_numpy(x).cuda() for x in batches[rank]), keep_mask=mask.astype(np.uint8))
            g = [np.concatenate([b[c] for b in batches]) for c in range(3)]
            ol = O.lightgcn_train_single_batch(p, st, O.edge_dropout(adj, mask, keep), g[0], g[1], g[2], L, decay,
                                               optimizer=optimizer, lr=lr)
            assert abs(loss - ol) <= 1e-5 * max(1, abs(ol)), (t, loss, ol)
        got = eng.gather_state()
        for k in p:
            if optimizer == "sgd":
                err = np.abs(got[k].astype(np.float64) - p[k]).max() / max(np.abs(p[k]).max(), 1e-30)
                assert err <= 1e-5, (k, err)
            else:
                assert np.abs(got[k].astype(np.float64) - p[k]).max() <= 2e-3 * lr + 1e-7, k
        # the device-drawn mask is the same on every rank (same seed, same generator state)
        m = eng.draw_keep_mask()
        ms = [torch.empty_like(m) for _ in range(world)]
        dist.all_gather(ms, m)
        assert all(torch.equal(ms[0], x) for x in ms)
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()[-1800:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,optimizer", [(1, "sgd"), (1, "adam"), (2, "sgd"), (2, "adam"), (4, "sgd")])
def test_sharded_lightgcn_matches_oracle_on_the_global_batch(world, optimizer):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = _free_port()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_lightgcn_worker, args=(r, world, port, optimizer, q)) for r in range(world)]
    saved = sys.path[:]
    sys.path[:] = [root] + [x for x in saved if x != root]
    try:
        for p in procs:
            p.start()
    finally:
        sys.path[:] = saved
    res = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
