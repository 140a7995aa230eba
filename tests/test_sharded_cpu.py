"""Host-side logic of the multi-GPU path on CPU: index rules against the oracle, and the
triple-routing exchange with the gloo backend at world_size 2 (the CUDA bucketing kernel
is replaced here by the oracle's stable bucketing -- same contract, checked bit-exactly
against the kernel in tests/test_sharded_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cf_oracle as O


def test_owner_rule_matches_oracle():
    from beta_recsys_b200 import sharded

    rows = np.arange(0, 1000, 7)
    for w in (1, 2, 4, 8):
        o1, l1 = sharded.owner_of(rows, w)
        o2, l2 = O.owner_of(rows, w)
        assert np.array_equal(o1, o2) and np.array_equal(l1, l2)
        assert sharded.local_rows(1001, w) == -(-1001 // w)


def test_shard_unshard_roundtrip():
    from beta_recsys_b200 import sharded

    rng = np.random.default_rng(0)
    for n in (1, 7, 64, 1001):
        full = rng.normal(size=(n, 3)).astype(np.float32)
        for w in (1, 2, 4, 8):
            parts = [sharded.shard_of(full, w, r) for r in range(w)]
            assert all(p.shape[0] == sharded.local_rows(n, w) for p in parts)
            assert np.array_equal(sharded.unshard(parts, n), full)
            for r in range(w):  # row g lives at parts[g % w][g // w]
                for g in range(r, n, w):
                    assert np.array_equal(parts[r][g // w], full[g])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _route_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from beta_recsys_b200 import sharded

        rng = np.random.default_rng(100 + rank)
        n = 257 + 13 * rank  # uneven batches
        u, p, ng = rng.integers(0, 1000, n), rng.integers(0, 50, n), rng.integers(0, 50, n)
        counts, ru, rp, rn, _ = O.route_triples(u, p, ng, world)  # stand-in for brs_route_triples
        send = torch.from_numpy(np.stack([ru, rp, rn], axis=1))
        recv, rc = sharded.all_to_all_v(send, counts.tolist())
        recv = recv.numpy()
        assert recv.shape[0] == sum(rc)
        assert np.all(recv[:, 0] % world == rank)  # every triple reached the owner of its user row
        # blocks arrive grouped by source rank, each in the source's original order
        everyone = [None] * world
        dist.all_gather_object(everyone, (u, p, ng))
        want = []
        for (su, sp, sn) in everyone:
            keep = su % world == rank
            want.append(np.stack([su[keep], sp[keep], sn[keep]], axis=1))
        assert np.array_equal(recv, np.concatenate(want, axis=0))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(400)
def test_triple_routing_exchange_gloo_world2():
    world, port = 2, _free_port()
    # spawned children unpickle the worker by module name: make sure they can import `tests.*` whatever
    # earlier tests did to sys.path / the working directory
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_route_worker, args=(r, world, port, q)) for r in range(world)]
    # children re-import this module by name from the parent's sys.path: keep the repo root first (an
    # earlier test may have put the reference checkout, which has its own `tests` package, in front)
    saved = sys.path[:]
    sys.path[:] = [root] + [x for x in saved if x != root]
    try:
        for p in procs:
            p.start()
    finally:
        sys.path[:] = saved
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


# --------------------------------------------------------------------------- #
# host-side index rules of the sharded NeuMF / LightGCN engines
# --------------------------------------------------------------------------- #
def test_lightgcn_row_partition_covers_every_node_once():
    from beta_recsys_b200.sharded_lightgcn import partition_rows

    for nu, ni, world in ((1501, 733, 2), (1_000_000, 100_000, 8), (7, 3, 4), (10, 10, 1)):
        seen = np.zeros(nu + ni, dtype=np.int64)
        for rank in range(world):
            parts, own_rows = partition_rows(nu, ni, world, rank)
            assert own_rows == sum(p["blk"] for p in parts)
            assert [p["own_off"] for p in parts] == [0, parts[0]["blk"]]
            for p in parts:
                assert p["off"] <= p["lo"] <= p["hi"] <= p["off"] + p["cnt"] and p["hi"] - p["lo"] <= p["blk"]
                seen[p["lo"]:p["hi"]] += 1
        assert (seen == 1).all()  # every user and item row has exactly one owner


def _neumf_route_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from beta_recsys_b200 import sharded
        from beta_recsys_b200.sharded_ncf import bucket_by_owner

        rng = np.random.default_rng(50 + rank)
        ids = torch.from_numpy(rng.integers(0, 1000, 257))
        order, counts, local = bucket_by_owner(ids, world)
        assert sorted(order.tolist()) == list(range(257)) and int(counts.sum()) == 257
        assert torch.equal((ids % world)[order], torch.sort(ids % world, stable=True).values)
        # owners receive local row numbers; serving rows and sending them back restores the batch order
        recv, rc = sharded.all_to_all_v(local, counts.tolist())
        table = torch.arange(1000 // world + 1, dtype=torch.float32).view(-1, 1) * world + rank  # local row r holds its global id
        served = table[recv]
        back, _ = sharded.all_to_all_v(served, rc)
        got = torch.empty(257, 1)
        got.index_copy_(0, order, back)
        assert torch.equal(got.view(-1).long(), ids)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(400)
def test_neumf_row_exchange_gloo_world2():
    world, port = 2, _free_port()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_neumf_route_worker, args=(r, world, port, q)) for r in range(world)]
    saved = sys.path[:]
    sys.path[:] = [root] + [x for x in saved if x != root]
    try:
        for p in procs:
            p.start()
    finally:
        sys.path[:] = saved
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
