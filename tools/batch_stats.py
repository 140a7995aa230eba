#!/usr/bin/env python
"""Row statistics of the benchmark's synthetic interaction stream (CPU, numpy/torch only): how many
distinct rows a batch touches, how many rows are touched exactly once, how concentrated the head is, and
how many bytes a row-sharded step has to move over NVLink.  The numbers behind DESIGN.md sections 5 and 8.

    python tools/batch_stats.py [--batches 8]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--items", type=int, default=100_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--batches", type=int, default=8)
    a = ap.parse_args()
    row_b = 4 * a.dim
    u, p, n = (t.numpy() for t in bench.make_batches(a.users, a.items, a.batch, a.batches * 8, bench.SEED, "cpu"))
    B = a.batch

    def stats(world):
        """per-rank batch B, global batch world*B (weak scaling): averages over a.batches steps"""
        acc = {}

        def add(k, v):
            acc.setdefault(k, []).append(float(v))

        for s in range(a.batches):
            for r in range(world):  # rank r of step s
                sl = slice((s * world + r) * B, (s * world + r + 1) * B)
                uu, pp, nn = u[sl], p[sl], n[sl]
                ii = np.concatenate([pp, nn])
                cu = np.unique(uu, return_counts=True)
                ci = np.unique(ii, return_counts=True)
                cp = np.unique(pp, return_counts=True)
                cn = np.unique(nn, return_counts=True)
                add("unique users", cu[0].size)
                add("unique pos items", cp[0].size)
                add("unique neg items", cn[0].size)
                add("unique items (pos+neg share a table)", ci[0].size)
                add("unique rows", cu[0].size + ci[0].size)
                single_u = cu[0][cu[1] == 1]
                single_i = ci[0][ci[1] == 1]
                su = np.isin(uu, single_u).sum()
                sp = np.isin(pp, single_i).sum()
                sn = np.isin(nn, single_i).sum()
                add("sample-rows whose row is touched exactly once (users)", su)
                add("sample-rows whose row is touched exactly once (pos)", sp)
                add("sample-rows whose row is touched exactly once (neg)", sn)
                add("share of the hottest user", cu[1].max() / B)
                add("share of the hottest pos item", cp[1].max() / B)
                # 32-sample tiles of the fused kernel: duplicates inside a tile
                tu = uu.reshape(-1, 32)
                tp = pp.reshape(-1, 32)
                add("distinct users per 32-sample tile", np.mean([np.unique(t).size for t in tu[:256]]))
                add("distinct pos items per 32-sample tile", np.mean([np.unique(t).size for t in tp[:256]]))
                if world > 1:
                    rem_s = ((uu % world != r).sum() + (pp % world != r).sum() + (nn % world != r).sum())
                    rem_u = (cu[0] % world != r).sum() + (ci[0] % world != r).sum()
                    add("NVLink MB in, direct (per-sample rows)", rem_s * row_b / 1e6)
                    add("NVLink MB in, staged = MB out, push (unique rows)", rem_u * row_b / 1e6)
        return {k: np.mean(v) for k, v in acc.items()}

    print("# synthetic stream statistics (bench.py generator, seed %d): %d users x %d items, D=%d, B=%d per rank\n"
          % (bench.SEED, a.users, a.items, a.dim, B))
    base = stats(1)
    print("| per rank and step | value |\n|---|---|")
    for k, v in base.items():
        print("| %s | %s |" % (k, ("%.3f" % v) if v < 50 else ("%.0f" % v)))
    tot = 3 * B
    once = sum(v for k, v in base.items() if k.startswith("sample-rows whose"))
    print("| sample-rows (3 x B) | %d |" % tot)
    print("| of them touched exactly once | %.0f (%.1f %%) |" % (once, 100 * once / tot))
    print("| bytes: per-sample rows / unique rows | %.1f MB / %.1f MB |"
          % (tot * row_b / 1e6, base["unique rows"] * row_b / 1e6))
    for w in (2, 4, 8):
        st = stats(w)
        print("\nworld %d: NVLink per rank and step: direct gathers %.1f MB in; staged pull %.1f MB in; push %.1f MB out"
              % (w, st["NVLink MB in, direct (per-sample rows)"], st["NVLink MB in, staged = MB out, push (unique rows)"],
                 st["NVLink MB in, staged = MB out, push (unique rows)"]))
        # owner imbalance if triples were routed to the user-row owner
        loads = []
        for s in range(a.batches):
            g = u[s * w * B:(s + 1) * w * B]
            loads.append(np.bincount(g % w, minlength=w).max() / (g.size / w))
        print("world %d: routing triples to the user-row owner: hottest rank receives %.2fx the mean batch" % (w, np.mean(loads)))


if __name__ == "__main__":
    main()
