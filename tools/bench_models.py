#!/usr/bin/env python
"""Step timings of the NCF-family and LightGCN engines on one B200 (diagnostics for DESIGN.md /
profiles/; the headline metric stays bench.py's MF-BPR line)."""
import io
import os
import sys
import time
from contextlib import redirect_stdout

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from beta_recsys_b200 import _lib  # noqa: E402
from beta_recsys_b200.engines import GMFEngine, LightGCNEngine, NeuMFEngine  # noqa: E402

dev = torch.device("cuda", 0)


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def ncf(kind, nu, ni, emb, nl, b, optimizer, mode, backend):
    _lib.load().brs_set_gemm_backend(backend)
    cfg = {"model": dict(model="ncf_end", device_str="cuda:0", n_users=nu, n_items=ni, emb_dim=emb, batch_size=b,
                         optimizer=optimizer, lr=1e-3, dropout=0.0, adam_mode=mode, mlp_config={"n_layers": nl}),
           "system": {"run_dir": "/tmp/x"}}
    with redirect_stdout(io.StringIO()):
        eng = (NeuMFEngine if kind == "neumf" else GMFEngine)(cfg)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    u = torch.randint(0, nu, (8 * b,), generator=g, device=dev)
    i = torch.randint(0, ni, (8 * b,), generator=g, device=dev)
    r = (torch.rand(8 * b, generator=g, device=dev) < 0.2).float()
    ms = timed(lambda: eng.train_batches(u, i, r), n=3, warm=1) / 8
    flops = 0.0
    if kind == "neumf":
        w = 2 * emb * 2 ** (nl - 1)
        flops = sum(2.0 * (w >> l) * (w >> (l + 1)) for l in range(nl)) * 3 * b
    print("%-6s U=%d I=%d emb=%d L=%d B=%d %s/%s gemm=%s | %.3f ms/step  %.1f M inter/s  tower %.1f TFLOP/s"
          % (kind, nu, ni, emb, nl, b, optimizer, mode, "tcgen05" if backend else "ffma", ms, b / ms / 1e3,
             flops / ms / 1e9), flush=True)
    del eng
    torch.cuda.empty_cache()


def lightgcn(nu, ni, n_edges, d, L, b, optimizer, rng_mode):
    import scipy.sparse as sp

    from oracle import cf_oracle as O  # graph construction only (host-side input preparation)

    rng = np.random.default_rng(0)
    pu = np.arange(1, nu + 1) ** -0.8
    pi = np.arange(1, ni + 1) ** -0.8
    eu = rng.choice(nu, n_edges, p=pu / pu.sum())
    ei = rng.choice(ni, n_edges, p=pi / pi.sum())
    adj = O.row_normalised_adj(nu, ni, eu, ei).tocoo()
    n = nu + ni
    tadj = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([adj.row, adj.col]).astype(np.int64)),
                                   torch.from_numpy(adj.data.astype(np.float32)), (n, n))
    cfg = {"model": dict(device_str="cuda:0", n_users=nu, n_items=ni, emb_dim=d, batch_size=b, optimizer=optimizer,
                         lr=0.05, regs=[1e-5], keep_pro=0.6, layer_size=[d] * L, norm_adj=tadj, dropout_rng=rng_mode),
           "system": {"run_dir": "/tmp/x"}}
    t0 = time.time()
    with redirect_stdout(io.StringIO()):
        eng = LightGCNEngine(cfg)
    eng.model.train()
    t_build = time.time() - t0
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    u = torch.randint(0, nu, (b,), generator=g, device=dev)
    i = torch.randint(0, ni, (b,), generator=g, device=dev)
    j = torch.randint(0, ni, (b,), generator=g, device=dev)
    ms = timed(lambda: eng.train_single_batch((u, i, j)), n=5, warm=2)
    nnz = adj.nnz
    bytes_spmm = nnz * (8 + 4 * d) + n * 4 * d
    print("lightgcn U=%d I=%d nnz=%d D=%d L=%d B=%d %s rng=%s | %.2f ms/step (CSR build %.1f s) | 2L SpMM bound %.1f GB "
          "-> %.0f GB/s if SpMM-only" % (nu, ni, nnz, d, L, b, optimizer, rng_mode, ms, t_build,
                                          2 * L * bytes_spmm / 1e9, 2 * L * bytes_spmm / ms / 1e6), flush=True)
    del eng
    torch.cuda.empty_cache()


if __name__ == "__main__":
    B = 65536
    if len(sys.argv) > 1:  # one short configuration, for an ncu launch list
        if sys.argv[1] == "neumf":
            ncf("neumf", 1_000_000, 100_000, 64, 3, B, "sgd", "dense", 1)
        elif sys.argv[1] == "lightgcn":
            lightgcn(200_000, 50_000, 4_000_000, 64, 3, B, "adam", "cuda")
        else:
            raise SystemExit("usage: bench_models.py [neumf|lightgcn]")
        sys.exit(0)
    for backend in (1, 0):
        ncf("neumf", 1_000_000, 100_000, 64, 3, B, "sgd", "dense", backend)
    ncf("neumf", 1_000_000, 100_000, 64, 3, B, "adam", "touched", 1)
    ncf("neumf", 1_000_000, 100_000, 64, 3, B, "adam", "dense", 1)
    ncf("gmf", 1_000_000, 100_000, 64, 1, B, "sgd", "dense", 1)
    lightgcn(200_000, 50_000, 4_000_000, 64, 3, B, "adam", "cuda")
    lightgcn(200_000, 50_000, 4_000_000, 64, 3, B, "adam", "cpu")
