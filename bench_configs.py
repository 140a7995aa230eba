"""bench.py --config 3 | 4 | 5: the other BASELINE.json configurations on ONE B200 (the driver's default
line stays config 2).  Each function returns the JSON line as a dict with the same contract as bench.py's:
metric / value / unit / ms_per_step, roofline (algorithmic bytes per step / CUDA-event step time against
the measured HBM peak), e2e (public API, pinned HOST inputs, H2D + D2H inside the timed region),
cpu_baseline (the reference's torch-CPU step from oracle/torch_port.py on a bounded sample), clocks.

Sizes follow SURVEY.md section 8d; where one GPU cannot hold BASELINE's 8-GPU shape the line says so in
config.workload.
"""
import io
import os
import time
from contextlib import redirect_stdout

import numpy as np
import torch

SEED = 2020


def _events(stream):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _time_steps(fn, steps, warmup, dev):
    """ms per call of fn() over `steps` calls after `warmup` (>= 3) untimed ones, CUDA events on the current stream."""
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = _events(None)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def _zipf(n, a, gen, device):
    import bench

    return bench.zipf_sampler(n, a, gen, device)


def _base_line(a, metric, unit, value, ms, config, roofline, clocks, e2e, launches, extra=None):
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": 1, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "roofline": roofline, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches}
    if extra:
        line.update(extra)
    return line


# --------------------------------------------------------------------------- #
# config 3: NeuMF (beta_rec/models/ncf.py), emb_dim 64, MLP 512 -> 256 -> 128 -> 64, Adam lr 1e-3
# --------------------------------------------------------------------------- #
def run_neumf(a, dev, sampler, peak):
    from beta_recsys_b200 import _lib
    from beta_recsys_b200.engines import NeuMFEngine

    hbm_peak, peak_src, tf_peak = peak
    nu, ni, emb, nl, b = a.users, a.items, 64, 3, a.batch
    mode = a.adam_mode
    _lib.load().brs_set_gemm_backend(1)
    cfg = {"model": dict(model="ncf_end", device_str=str(dev), n_users=nu, n_items=ni, emb_dim=emb, batch_size=b,
                         optimizer="adam", lr=1e-3, dropout=0.0, adam_mode=mode, mlp_config={"n_layers": nl}),
           "system": {"run_dir": "/tmp/brs_bench"}}
    torch.manual_seed(SEED)
    with redirect_stdout(io.StringIO()):
        eng = NeuMFEngine(cfg)
    g = torch.Generator(device=dev)
    g.manual_seed(SEED)
    nb = 16
    u = _zipf(nu, 1.05, g, dev)(nb * b)
    i = _zipf(ni, 1.05, g, dev)(nb * b)
    r = (torch.rand(nb * b, generator=g, device=dev) < 0.2).float()  # 1 positive : 4 negatives (ncf_default.json:35)
    steps = max(nb, (a.steps // nb) * nb) if a.steps >= nb else nb
    steps = min(steps, 256)
    t0 = time.time()
    ms = _time_steps(lambda: eng.train_batches(u, i, r), steps // nb, max(1, a.warmup // nb), dev) / nb
    while time.time() - t0 < 1.2:
        eng.train_batches(u, i, r)
    clocks = sampler.summary(t0, time.time())
    w = 2 * emb * 2 ** (nl - 1)  # 512
    flops = sum(2.0 * (w >> l) * (w >> (l + 1)) for l in range(nl)) + 2.0 * (2 * emb)  # forward, per interaction
    flops_step = 3.0 * flops * b  # forward + dgrad + wgrad
    row_floats = 2 * ((w // 2) + emb)  # user + item rows of the MLP and MF tables
    # touched-rows Adam: read w, m, v and write w, m, v of every gathered row; + 2 ids + rating
    alg = (6 * 4 * row_floats + 20) * b if mode == "touched" else None
    step_s = ms * 1e-3
    roof = {"bound": "hbm", "kernel": "NeuMF step (ncf_gather / tcgen05 tower / ncf_scatter + row Adam)", "unit": "GB/s",
            "peak": hbm_peak, "peak_source": peak_src, "traffic": None,
            "tensor": {"achieved": flops_step / step_s / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                       "frac": flops_step / step_s / 1e12 / tf_peak,
                       "note": "tower GEMM flops (fwd + dgrad + wgrad, 3xTF32-split on tcgen05) / WHOLE step time against the "
                               "measured bf16 peak: a lower bound on the tower's own rate"}}
    if alg is not None:
        roof.update(achieved=alg / step_s / 1e9, frac=alg / step_s / 1e9 / hbm_peak, algorithmic_bytes_per_launch=alg,
                    note="(6 x 4 B x %d row floats + 20 B) per interaction x batch / CUDA-event step time" % row_floats)
    else:
        dense = 24.0 * (nu + ni) * (w // 2 + emb)
        roof.update(achieved=dense / step_s / 1e9, frac=dense / step_s / 1e9 / hbm_peak, algorithmic_bytes_per_launch=dense,
                    note="reference-exact dense Adam: 24 B x every table element per step dominates")
    # e2e: public per-batch API with pinned host tensors
    hu, hi, hr = u[: 8 * b].cpu().pin_memory(), i[: 8 * b].cpu().pin_memory(), r[: 8 * b].cpu().pin_memory()
    k = [0]

    def host_step():
        s = slice((k[0] % 8) * b, (k[0] % 8 + 1) * b)
        k[0] += 1
        eng.train_single_batch(hu[s], hi[s], hr[s])

    n_e2e = min(a.steps, 48)
    ms_e2e = _time_steps(host_step, n_e2e, 3, dev)
    e2e = {"value": b / (ms_e2e * 1e-3), "unit": "interactions/s", "h2d_bytes_per_step": 20 * b, "d2h_bytes_per_step": 16,
           "steps": n_e2e, "api": "NeuMFEngine.train_single_batch(users, items, ratings) with pinned host tensors (one host "
                                  "round trip per step, like the reference's .item())"}
    config = {"workload": "configs[2] on ONE GPU: NeuMF %dM users x %dM items, emb_dim=64, MLP=[256,128,64], batch=%d "
                          "(BASELINE quotes it on 8 GPUs row-sharded; all four tables fit one B200 here)"
                          % (nu // 1_000_000, ni // 1_000_000, b),
              "n_users": nu, "n_items": ni, "emb_dim": emb, "mlp_layers": [w, w // 2, w // 4, w // 8], "batch_per_gpu": b,
              "optimizer": "adam", "optimizer_mode": mode, "lr": 1e-3, "dropout": 0.0,
              "index_distribution": "user,item ~ Zipf(1.05) on permuted ids; ratings 1:4", "prebuilt_batches": nb,
              "l2_policy": "inputs larger than L2: %.1f GB of tables" % ((nu + ni) * (w // 2 + emb) * 4 / 1e9)}
    return _base_line(a, "BCE interactions/sec (NeuMF)", "interactions/s", b / step_s, ms, config, roof, clocks, e2e,
                      None, {"cpu_baseline_fn": "neumf"})


def run_neumf_sharded(a, rank, world, local, dev, sampler, peak):
    """configs[2] as BASELINE states it: NeuMF with the four embedding tables row-sharded over the ranks
    (beta_recsys_b200/sharded_ncf.py), every rank feeding its own batch; weak scaling."""
    import torch.distributed as dist

    from beta_recsys_b200.sharded_ncf import ShardedNeuMFEngine

    hbm_peak, peak_src, tf_peak = peak
    nu, ni, emb, nl, b = a.users, a.items, 64, 3, a.batch
    cfg = {"model": dict(model="ncf_end", device_str=str(dev), n_users=nu, n_items=ni, emb_dim=emb, batch_size=b,
                         optimizer="adam", lr=1e-3, dropout=0.0, adam_mode=a.adam_mode, mlp_config={"n_layers": nl}),
           "system": {"run_dir": "/tmp/brs_bench"}}
    with redirect_stdout(io.StringIO()):
        eng = ShardedNeuMFEngine(cfg)
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + rank)
    nb = 8
    u = _zipf(nu, 1.05, g, dev)(nb * b)
    i = _zipf(ni, 1.05, g, dev)(nb * b)
    r = (torch.rand(nb * b, generator=g, device=dev) < 0.2).float()
    k = [0]

    def step():
        s = slice((k[0] % nb) * b, (k[0] % nb + 1) * b)
        k[0] += 1
        return eng.train_single_batch(u[s], i[s], r[s])

    steps = min(a.steps, 40)
    for _ in range(3):
        step()
    dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = _events(None)
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    while time.time() - t0 < 1.2:
        step()
    clocks = sampler.summary(t0, time.time())
    if rank != 0:
        return None
    w = 2 * emb * 2 ** (nl - 1)
    row_floats = 2 * (w // 2 + emb)
    nvl = 2.0 * (world - 1) / world * b * row_floats * 4  # rows in + gradient rows out, per direction per rank
    step_s = ms * 1e-3
    roof = {"bound": "nvlink", "kernel": "NCCL all-to-all of embedding rows (forward) and gradient rows (backward)",
            "achieved": nvl / step_s / 1e9, "peak": 770.0, "unit": "GB/s", "frac": nvl / step_s / 1e9 / 770.0,
            "peak_source": "B200_PROFILING.md measured peer copy, per direction per GPU", "traffic": None,
            "note": "bytes that must cross NVLink per rank, step and direction (remote fraction x batch x %d row floats x 4 B, "
                    "forward + backward) / whole-step time" % row_floats}
    config = {"workload": "configs[2]: NeuMF %dM users x %dM items, emb_dim=64, MLP=[256,128,64], embeddings row-sharded over %d "
                          "B200, batch=%d per rank" % (nu // 1_000_000, ni // 1_000_000, world, b),
              "n_users": nu, "n_items": ni, "emb_dim": emb, "batch_per_gpu": b, "global_batch": b * world, "optimizer": "adam",
              "optimizer_mode": a.adam_mode, "lr": 1e-3, "parallelism": "embedding rows sharded x%d (owner = row mod N), tower replicated"
              % world}
    line = _base_line(a, "BCE interactions/sec (NeuMF)", "interactions/s", b * world / step_s, ms, config, roof, clocks, None, None,
                      {"final_loss": loss})
    line["n_gpus"] = world
    line["steps"] = steps
    return line


# --------------------------------------------------------------------------- #
# config 4: LightGCN (beta_rec/models/lightgcn.py), 3 layers, dim 64, keep_pro 0.6
# --------------------------------------------------------------------------- #
def run_lightgcn(a, dev, sampler, peak):
    from beta_recsys_b200 import graph
    from beta_recsys_b200.engines import LightGCNEngine

    hbm_peak, peak_src, _ = peak
    nu, ni, d, L, b = a.users, a.items, 64, 3, a.batch
    n_edges = a.edges
    g = torch.Generator(device=dev)
    g.manual_seed(SEED)
    eu = _zipf(nu, 0.8, g, dev)(n_edges)
    ei = _zipf(ni, 0.8, g, dev)(n_edges)
    t0 = time.time()
    adj = graph.build_norm_adj(eu, ei, nu, ni, device=dev)
    torch.cuda.synchronize(dev)
    t_adj = time.time() - t0
    cfg = {"model": dict(device_str=str(dev), n_users=nu, n_items=ni, emb_dim=d, batch_size=b, optimizer="adam", lr=0.05,
                         regs=[1e-5], keep_pro=0.6, layer_size=[d] * L, norm_adj=adj, dropout_rng=a.dropout_rng),
           "system": {"run_dir": "/tmp/brs_bench"}}
    torch.manual_seed(SEED)
    with redirect_stdout(io.StringIO()):
        eng = LightGCNEngine(cfg)
    eng.model.train()
    u = torch.randint(0, nu, (8 * b,), generator=g, device=dev)
    i = torch.randint(0, ni, (8 * b,), generator=g, device=dev)
    j = torch.randint(0, ni, (8 * b,), generator=g, device=dev)
    k = [0]

    def step():
        s = slice((k[0] % 8) * b, (k[0] % 8 + 1) * b)
        k[0] += 1
        eng.train_single_batch((u[s], i[s], j[s]))

    steps = min(a.steps, 40)
    t0 = time.time()
    ms = _time_steps(step, steps, min(a.warmup, 5), dev)
    while time.time() - t0 < 1.2:
        step()
    clocks = sampler.summary(t0, time.time())
    n, nnz = nu + ni, adj.nnz
    spmm = nnz * (8 + 4 * d) + n * 4 * d  # col + val + one row read per non-zero (no reuse) + row write
    alg = 2 * L * spmm
    step_s = ms * 1e-3
    roof = {"bound": "hbm", "kernel": "spmm_csr_kernel x %d (forward + transposed backward)" % (2 * L),
            "achieved": alg / step_s / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": alg / step_s / 1e9 / hbm_peak,
            "peak_source": peak_src, "traffic": None, "algorithmic_bytes_per_launch": spmm,
            "note": "no-reuse upper bound of SURVEY.md 8d (nnz x (8 + 4D) + N x 4D per SpMM) x 2L / WHOLE step time; the step "
                    "also holds the mask draw (%s rng), tail and Adam over all N rows" % a.dropout_rng}
    hu, hi, hj = u.cpu().pin_memory(), i.cpu().pin_memory(), j.cpu().pin_memory()

    def host_step():
        s = slice((k[0] % 8) * b, (k[0] % 8 + 1) * b)
        k[0] += 1
        eng.train_single_batch((hu[s], hi[s], hj[s]))

    n_e2e = min(a.steps, 16)
    ms_e2e = _time_steps(host_step, n_e2e, 3, dev)
    e2e = {"value": b / (ms_e2e * 1e-3), "unit": "interactions/s", "h2d_bytes_per_step": 24 * b, "d2h_bytes_per_step": 16,
           "steps": n_e2e, "api": "LightGCNEngine.train_single_batch((users, pos, neg)) with pinned host LongTensors"}
    config = {"workload": "configs[3] on ONE GPU: LightGCN %dM x %dk, 3 layers, dim=64, %d interactions (nnz(A_hat) = %d), "
                          "batch=%d (BASELINE quotes it on 8 GPUs)" % (nu // 1_000_000, ni // 1000, n_edges, nnz, b),
              "n_users": nu, "n_items": ni, "dim": d, "layers": L, "nnz": nnz, "batch_per_gpu": b, "optimizer": "adam",
              "lr": 0.05, "keep_pro": 0.6, "dropout_rng": a.dropout_rng, "adjacency_build_s": t_adj,
              "l2_policy": "inputs larger than L2: %.2f GB of CSR + %.2f GB of embeddings per layer"
                           % (nnz * 12 / 1e9, n * d * 4 / 1e9)}
    return _base_line(a, "BPR interactions/sec (LightGCN, whole-graph propagate per batch)", "interactions/s", b / step_s, ms,
                      config, roof, clocks, e2e, None, {"cpu_baseline_fn": "lightgcn", "graph_steps_per_s": 1.0 / step_s})


def run_lightgcn_sharded(a, rank, world, local, dev, sampler, peak):
    """configs[3] as BASELINE states it: LightGCN with the node rows partitioned over the ranks
    (beta_recsys_b200/sharded_lightgcn.py), every rank feeding its own batch."""
    import torch.distributed as dist

    from beta_recsys_b200 import graph
    from beta_recsys_b200.sharded_lightgcn import ShardedLightGCNEngine

    hbm_peak, peak_src, _ = peak
    nu, ni, d, L, b = a.users, a.items, 64, 3, a.batch
    g = torch.Generator(device=dev)
    g.manual_seed(SEED)  # the same graph on every rank
    eu = _zipf(nu, 0.8, g, dev)(a.edges)
    ei = _zipf(ni, 0.8, g, dev)(a.edges)
    adj = graph.build_norm_adj(eu, ei, nu, ni, device=dev)
    cfg = {"model": dict(device_str=str(dev), n_users=nu, n_items=ni, emb_dim=d, batch_size=b, optimizer="adam", lr=0.05,
                         regs=[1e-5], keep_pro=0.6, layer_size=[d] * L, norm_adj=adj)}
    eng = ShardedLightGCNEngine(cfg)
    g.manual_seed(SEED + 1 + rank)
    u = torch.randint(0, nu, (8 * b,), generator=g, device=dev)
    i = torch.randint(0, ni, (8 * b,), generator=g, device=dev)
    j = torch.randint(0, ni, (8 * b,), generator=g, device=dev)
    k = [0]

    def step():
        s = slice((k[0] % 8) * b, (k[0] % 8 + 1) * b)
        k[0] += 1
        return eng.train_single_batch((u[s], i[s], j[s]))

    steps = min(a.steps, 30)
    for _ in range(3):
        step()
    dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = _events(None)
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    while time.time() - t0 < 1.2:
        step()
    clocks = sampler.summary(t0, time.time())
    if rank != 0:
        return None
    n, nnz = nu + ni, adj.nnz
    step_s = ms * 1e-3
    coll = (2 * L + 1) * (world - 1) / world * n * d * 4  # bytes a rank receives per step (all-gathers + reduce-scatter)
    roof = {"bound": "nvlink", "kernel": "NCCL all-gather of E(l) / G(l) per layer + reduce-scatter of d",
            "achieved": coll / step_s / 1e9, "peak": 770.0, "unit": "GB/s", "frac": coll / step_s / 1e9 / 770.0,
            "peak_source": "B200_PROFILING.md measured peer copy, per direction per GPU", "traffic": None,
            "note": "(2L + 1) x (N-1)/N x n_nodes x D x 4 B received per rank and step / whole-step time; the block SpMMs "
                    "(1/N of %.1f GB each by the no-reuse bound) run between the collectives" % ((nnz * (8 + 4 * d) + n * 4 * d) / 1e9)}
    config = {"workload": "configs[3]: LightGCN %dM x %dk, 3 layers, dim=64, nnz(A_hat) = %d, node rows partitioned over %d B200, "
                          "batch=%d per rank" % (nu // 1_000_000, ni // 1000, nnz, world, b),
              "n_users": nu, "n_items": ni, "dim": d, "layers": L, "nnz": nnz, "batch_per_gpu": b, "global_batch": b * world,
              "optimizer": "adam", "lr": 0.05, "keep_pro": 0.6, "dropout_rng": "cuda (same mask on every rank)",
              "parallelism": "1-D row partition x%d, all-gather per layer" % world}
    line = _base_line(a, "BPR interactions/sec (LightGCN, whole-graph propagate per batch)", "interactions/s", b * world / step_s, ms,
                      config, roof, clocks, None, None, {"final_loss": loss, "graph_steps_per_s": 1.0 / step_s})
    line["n_gpus"] = world
    line["steps"] = steps
    return line


# --------------------------------------------------------------------------- #
# config 5: embedding gather / scatter-add / gather+SGD microbench
# --------------------------------------------------------------------------- #
def run_microbench(a, dev, sampler, peak):
    from beta_recsys_b200 import _lib

    hbm_peak, peak_src, _ = peak
    lib = _lib.load()
    n_rows, n_idx = a.rows, 1 << 20
    st = torch.cuda.current_stream(dev).cuda_stream
    g = torch.Generator(device=dev)
    g.manual_seed(SEED)
    idx_sets = [_zipf(n_rows, 1.05, g, dev)(n_idx) for _ in range(4)]
    results, t_first = {}, time.time()
    for d in (32, 64, 128, 256):
        table = torch.empty((n_rows, d), dtype=torch.float32, device=dev).normal_(0, 0.1)
        buf = torch.empty((n_idx, d), dtype=torch.float32, device=dev).normal_(0, 0.01)
        k = [0]

        def op(which):
            idx = idx_sets[k[0] % 4]
            k[0] += 1
            if which == "gather":
                rc = lib.brs_gather(_lib.ptr(table), n_rows, d, _lib.ptr(idx), n_idx, _lib.ptr(buf), st)
            elif which == "scatter_add":
                rc = lib.brs_scatter_add(_lib.ptr(table), n_rows, d, _lib.ptr(idx), n_idx, _lib.ptr(buf), 1e-6, st)
            else:
                rc = lib.brs_gather_sgd_update(_lib.ptr(table), n_rows, d, _lib.ptr(idx), n_idx, 1e-6, _lib.ptr(buf), st)
            _lib.check(rc, which)

        steps = min(a.steps, 50)
        for which, per_idx in (("gather", 8 + 4 * d), ("scatter_add", 8 + 8 * d), ("gather_sgd", 8 + 8 * d)):
            ms = _time_steps(lambda: op(which), steps, 3, dev)
            gbs = per_idx * n_idx / (ms * 1e-3) / 1e9
            results["%s_d%d" % (which, d)] = {"ms": ms, "GB/s": gbs, "frac": gbs / hbm_peak,
                                              "algorithmic_bytes": per_idx * n_idx}
        del table, buf
        torch.cuda.empty_cache()
    clocks = sampler.summary(t_first, time.time())
    head = results["gather_d128"]
    roof = {"bound": "hbm", "kernel": "rows_op_kernel<GATHER> at D=128 (the other 11 points are in `sweep`)",
            "achieved": head["GB/s"], "peak": hbm_peak, "unit": "GB/s", "frac": head["frac"], "peak_source": peak_src,
            "traffic": None, "algorithmic_bytes_per_launch": head["algorithmic_bytes"],
            "note": "SURVEY.md 8d: gather 8 + 4D B per index, scatter-add and gather+SGD 8 + 8D B per index; duplicates under "
                    "Zipf(1.05) hit L2, so achieved can exceed the DRAM peak"}
    # e2e: indices arrive from pinned host memory, the gathered rows go back to the host
    d = 128
    table = torch.empty((n_rows, d), dtype=torch.float32, device=dev).normal_(0, 0.1)
    out = torch.empty((n_idx, d), dtype=torch.float32, device=dev)
    h_idx = idx_sets[0].cpu().pin_memory()
    h_out = torch.empty((n_idx, d), dtype=torch.float32).pin_memory()

    def host_gather():
        di = h_idx.to(dev, non_blocking=True)
        _lib.check(lib.brs_gather(_lib.ptr(table), n_rows, d, _lib.ptr(di), n_idx, _lib.ptr(out), st), "gather")
        h_out.copy_(out, non_blocking=True)

    ms_e2e = _time_steps(host_gather, 10, 3, dev)
    e2e = {"value": (8 + 4 * d) * n_idx / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 8 * n_idx,
           "d2h_bytes_per_step": 4 * d * n_idx, "steps": 10,
           "api": "brs_gather at D=128 with indices from pinned host memory and the gathered rows copied back to the host"}
    config = {"workload": "configs[4] on ONE GPU: embedding gather / scatter-add / gather+SGD, %dM rows x dim in {32,64,128,256}, "
                          "Zipf(1.05) indices, 2^20 indices per call" % (n_rows // 1_000_000),
              "n_rows": n_rows, "indices_per_call": n_idx, "index_distribution": "Zipf(1.05) on permuted ids",
              "l2_policy": "tables of %.1f-%.1f GB, 4 index sets cycled" % (n_rows * 32 * 4 / 1e9, n_rows * 256 * 4 / 1e9)}
    line = _base_line(a, "embedding gather HBM GB/s (D=128)", "GB/s", head["GB/s"], head["ms"], config, roof, clocks, e2e,
                      12 * min(a.steps, 50), {"sweep": results})
    return line


# --------------------------------------------------------------------------- #
# CPU baselines (oracle/torch_port.py: the reference's torch-CPU step) -- bench.py's cpu_baseline leg
# --------------------------------------------------------------------------- #
def cpu_neumf(a, budget_s=20.0):
    from oracle.torch_port import NeuMFPort  # cpu_baseline leg: the one place the bench may run oracle/

    torch.set_num_threads(os.cpu_count() or 1)
    nu, ni = min(a.users, 1_000_000), min(a.items, 100_000)  # dense Adam over the full 14 GB of tables does not fit the sample budget
    emb, nl, b = 64, 3, a.batch
    w = 2 * emb * 2 ** (nl - 1)
    g = torch.Generator().manual_seed(SEED)
    st = {"embedding_user_mlp.weight": torch.randn(nu, w // 2, generator=g) * 0.01,
          "embedding_item_mlp.weight": torch.randn(ni, w // 2, generator=g) * 0.01,
          "embedding_user_mf.weight": torch.randn(nu, emb, generator=g) * 0.01,
          "embedding_item_mf.weight": torch.randn(ni, emb, generator=g) * 0.01,
          "affine_output.weight": torch.randn(1, w // 8 + emb, generator=g) * 0.1, "affine_output.bias": torch.zeros(1)}
    for l in range(nl):
        st["fc_layers.%d.weight" % (3 * l + 1)] = torch.randn(w >> (l + 1), w >> l, generator=g) * 0.05
        st["fc_layers.%d.bias" % (3 * l + 1)] = torch.zeros(w >> (l + 1))
    port = NeuMFPort(st, nl, "adam", 1e-3)
    u = torch.randint(0, nu, (b,), generator=g)
    i = torch.randint(0, ni, (b,), generator=g)
    r = (torch.rand(b, generator=g) < 0.2).float()
    t0 = time.time()
    port.train_single_batch(u, i, r)
    first = time.time() - t0
    n = int(max(2, min(20, budget_s / max(first, 1e-3))))
    ts = []
    for _ in range(n):
        t0 = time.time()
        port.train_single_batch(u, i, r)
        ts.append(time.time() - t0)
    med = float(np.median(ts))
    return {"value": b / med, "unit": "interactions/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d timed + 1 warm-up NeuMF steps of the reference's torch-CPU path (oracle/torch_port.py, dense Adam) at the "
                      "SCALED size %d x %d (the 10M x 1M tables + Adam state exceed the sample budget), batch %d, median %.0f ms/step"
                      % (n, nu, ni, b, med * 1e3)}


def cpu_lightgcn(a, budget_s=20.0):
    from oracle.torch_port import LightGCNPort  # cpu_baseline leg

    torch.set_num_threads(os.cpu_count() or 1)
    nu, ni, d, L, b = min(a.users, 200_000), min(a.items, 50_000), 64, 3, a.batch
    e = min(a.edges, 4_000_000)
    g = torch.Generator().manual_seed(SEED)
    eu, ei = torch.randint(0, nu, (e,), generator=g), torch.randint(0, ni, (e,), generator=g)
    n = nu + ni
    rows = torch.cat([eu, ei + nu, torch.arange(n)])
    cols = torch.cat([ei + nu, eu, torch.arange(n)])
    adj = torch.sparse_coo_tensor(torch.stack([rows, cols]), torch.ones(rows.numel()), (n, n)).coalesce()
    deg = torch.zeros(n).index_add_(0, adj.indices()[0], torch.ones(adj._nnz()))
    adj = torch.sparse_coo_tensor(adj.indices(), 1.0 / deg[adj.indices()[0]], (n, n)).coalesce()
    st = {"user_embedding.weight": torch.randn(nu, d, generator=g) * 0.05, "item_embedding.weight": torch.randn(ni, d, generator=g) * 0.05}
    port = LightGCNPort(st, adj, L, 1e-5, 0.6, "adam", 0.05)
    batch = (torch.randint(0, nu, (b,), generator=g), torch.randint(0, ni, (b,), generator=g), torch.randint(0, ni, (b,), generator=g))
    mask = (torch.rand(adj._nnz(), generator=g) + 0.6).int().bool()
    t0 = time.time()
    port.train_single_batch(batch, mask)
    first = time.time() - t0
    k = int(max(1, min(10, budget_s / max(first, 1e-3))))
    ts = []
    for _ in range(k):
        t0 = time.time()
        port.train_single_batch(batch, mask)
        ts.append(time.time() - t0)
    med = float(np.median(ts))
    return {"value": b / med, "unit": "interactions/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d timed + 1 warm-up LightGCN steps of the reference's torch-CPU path (oracle/torch_port.py) at the SCALED "
                      "size %d x %d, nnz %d, batch %d, median %.0f ms/step" % (k, nu, ni, adj._nnz(), b, med * 1e3)}


def cpu_gather(a, budget_s=10.0):
    torch.set_num_threads(os.cpu_count() or 1)
    n_rows, d, n_idx = min(a.rows, 20_000_000), 128, 1 << 20
    g = torch.Generator().manual_seed(SEED)
    table = torch.randn(n_rows, d, generator=g)
    idx = torch.randint(0, n_rows, (n_idx,), generator=g)
    torch.nn.functional.embedding(idx, table)
    ts = []
    t_end = time.time() + budget_s
    while time.time() < t_end and len(ts) < 20:
        t0 = time.time()
        torch.nn.functional.embedding(idx, table)
        ts.append(time.time() - t0)
    med = float(np.median(ts))
    return {"value": (8 + 4 * d) * n_idx / med / 1e9, "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d calls of torch.nn.functional.embedding (the reference's gather, models/mf.py:39-40) on the host, "
                      "%dM x 128 table (scaled: host RAM), 2^20 indices, median %.1f ms" % (len(ts), n_rows // 1_000_000, med * 1e3)}
