"""Per-kernel time of the row-sharded MF step at N>1 (torchrun), via CUPTI activity records.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/shard_phases.py [--steps 100]

ncu cannot attach to a multi-rank job whose kernels spin on peers, so this uses torch.profiler (kineto/CUPTI
kernel activity only): the numbers include profiler overhead between launches but per-kernel durations are
the device's own.  Prints rank 0's table.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--items", type=int, default=100_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=65536)
    a = ap.parse_args()
    import torch.distributed as dist
    from beta_recsys_b200.sharded import ShardedMFEngine

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = {"model": dict(device_str="cuda:%d" % local, n_users=a.users, n_items=a.items, emb_dim=a.dim,
                         batch_size=a.batch, optimizer="sgd", lr=0.05, loss="bpr")}
    eng = ShardedMFEngine(cfg, route="none")
    nb = 64
    users, pos, neg = bench.make_batches(a.users, a.items, a.batch, nb, bench.SEED + rank, dev)
    eng.train_batches(users, pos, neg)
    dist.barrier()
    torch.cuda.synchronize(dev)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        done = 0
        while done < a.steps:
            k = min(nb, a.steps - done)
            eng.train_batches(users[: k * a.batch], pos[: k * a.batch], neg[: k * a.batch])
            done += k
        torch.cuda.synchronize(dev)
    dist.barrier()
    if rank == 0:
        rows = []
        for e in prof.key_averages():
            t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
            if t:
                rows.append((t / a.steps, e.count / a.steps, e.key[:110]))
        rows.sort(reverse=True)
        print("world=%d  us/step  launches/step  kernel" % world)
        for t, c, k in rows:
            print("%9.1f  %5.2f  %s" % (t, c, k))
        print("sum %.1f us/step" % sum(r[0] for r in rows))
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
