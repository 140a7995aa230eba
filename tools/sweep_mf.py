#!/usr/bin/env python
"""Per-kernel timing sweep of the MF step over table sizes / index skew / batch / dim
(diagnostics for profiles/; NOT a bench value).  Prints one line per configuration."""
import ctypes
import io
import os
import sys
from contextlib import redirect_stdout

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from beta_recsys_b200 import _lib  # noqa: E402
from beta_recsys_b200.engines import MFEngine  # noqa: E402


def ids(n, size, a, gen, dev):
    if a <= 0:
        return torch.randint(0, n, (size,), generator=gen, device=dev, dtype=torch.int64)
    ranks = torch.arange(1, n + 1, dtype=torch.float64, device=dev)
    cdf = torch.cumsum(ranks.pow(-a), 0)
    cdf = (cdf / cdf[-1]).float()
    perm = torch.randperm(n, generator=gen, device=dev)
    return perm[torch.searchsorted(cdf, torch.rand(size, generator=gen, device=dev)).clamp_(max=n - 1)]


def run(nu, ni, d, b, a, nb=32, reps=40, sort_users=False, variant=0, pol=(0, 0, 0)):
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    lib.brs_debug_set_mf_variant(variant)
    lib.brs_debug_set_l2_policy(*pol)
    cfg = {"model": dict(device_str="cuda:0", n_users=nu, n_items=ni, emb_dim=d, batch_size=b, optimizer="sgd",
                         lr=0.05, loss="bpr"), "system": {"run_dir": "/tmp/x"}}
    with redirect_stdout(io.StringIO()):
        eng = MFEngine(cfg)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    u, p, n = ids(nu, b * nb, a, gen, dev), ids(ni, b * nb, a, gen, dev), ids(ni, b * nb, 0, gen, dev)
    if sort_users:
        u = u.view(nb, b).sort(dim=1).values.reshape(-1).contiguous()
    st = torch.cuda.current_stream(dev)
    out = torch.empty(4, device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(reps)]
    uniq = []
    for k in range(reps + 5):
        off = (k % nb) * b
        pu, pp, pn = u[off:].data_ptr(), p[off:].data_ptr(), n[off:].data_ptr()
        e = ev[k - 5] if k >= 5 else None
        if e: e[0].record(st)
        _lib.check(lib.brs_mf_bpr_prepare(eng._cmodel, pu, pp, pn, b, st.cuda_stream))
        if e: e[1].record(st)
        if k == 5:
            uniq = (int(eng._user.count.item()), int(eng._item.count.item()))
        _lib.check(lib.brs_mf_bpr_fwd_bwd_prepared(eng._cmodel, pu, pp, pn, b, 0.0, st.cuda_stream))
        if e: e[2].record(st)
        _lib.check(lib.brs_mf_apply(eng._cmodel, eng.optimizer.desc, b, out.data_ptr(), st.cuda_stream))
        if e: e[3].record(st)
    torch.cuda.synchronize()
    t = [float(np.median([e[i].elapsed_time(e[i + 1]) for e in ev])) * 1e3 for i in range(3)]
    alg = (24 * d + 48) * b
    print("pol=%s var=%2d U=%8d I=%7d D=%3d B=%6d zipf=%.2f sortu=%d | uniq u/i %6d/%6d | prep %6.1f fwd %6.1f apply %6.1f us | "
          "fwd %.0f GB/s  step %.0f M inter/s" % ("".join(map(str, pol)), variant, nu, ni, d, b, a, sort_users, uniq[0], uniq[1], t[0], t[1], t[2],
                                                   alg / t[1] / 1e3, b / sum(t)), flush=True)
    del eng
    torch.cuda.empty_cache()


if __name__ == "__main__":
    B = 65536
    if len(sys.argv) > 1 and sys.argv[1] == "l2":
        for pol in ((0, 0, 0), (0, 2, 0), (1, 2, 0), (1, 2, 1), (0, 2, 1), (1, 0, 0), (1, 2, 2), (2, 2, 0)):
            run(1_000_000, 100_000, 128, B, 1.05, pol=pol)
        for pol in ((0, 0, 0), (1, 2, 1)):
            run(1_000_000, 100_000, 128, B, 0.0, pol=pol)
            run(4_000_000, 1_000_000, 128, B, 1.05, pol=pol)
            run(1_000_000, 100_000, 64, B, 1.05, pol=pol)
            run(1_000_000, 100_000, 256, B, 1.05, pol=pol)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        for v in range(0, 11):
            run(1_000_000, 100_000, 128, B, 1.05, variant=v)
        for v in (0, 2, 3, 4, 6, 8):
            run(20_000, 20_000, 128, B, 0.0, variant=v)
        sys.exit(0)
    run(1_000_000, 100_000, 128, B, 1.05)
    run(1_000_000, 100_000, 128, B, 0.0)
    run(100_000, 100_000, 128, B, 1.05)
    run(20_000, 20_000, 128, B, 0.0)
    run(4_000_000, 1_000_000, 128, B, 0.0)
    run(1_000_000, 100_000, 128, 4 * B, 1.05, nb=8)
    run(1_000_000, 100_000, 128, B // 4, 1.05)
    run(1_000_000, 100_000, 64, B, 1.05)
    run(1_000_000, 100_000, 256, B, 1.05)
    run(1_000_000, 100_000, 32, B, 1.05)
