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
