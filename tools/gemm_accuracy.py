#!/usr/bin/env python
"""Error of the tensor-core (3xTF32) and FFMA Linear kernels against float64, and their timings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from beta_recsys_b200 import _lib  # noqa: E402

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(0)
for (m, n, k) in [(4096, 64, 128), (4096, 128, 256), (4096, 256, 512), (65536, 256, 512), (65536, 128, 256), (65536, 64, 128)]:
    x = rng.normal(0, 1, (m, k)).astype(np.float32)
    w = (rng.normal(0, 1, (n, k)) / np.sqrt(k)).astype(np.float32)
    b = np.zeros(n, np.float32)
    tx, tw, tb = (torch.from_numpy(a).cuda() for a in (x, w, b))
    want = None
    if m <= 4096:
        want = x.astype(np.float64) @ w.T.astype(np.float64)
    res = {}
    for name in ("tc", "simt"):
        ty = torch.empty((m, n), device="cuda")
        def run():
            if name == "tc":
                _lib.check(lib.brs_mlp_fwd_tc(tx.data_ptr(), tw.data_ptr(), tb.data_ptr(), ty.data_ptr(), None, m, n, k, 0, st))
            else:
                _lib.check(lib.brs_mlp_fwd(tx.data_ptr(), tw.data_ptr(), tb.data_ptr(), ty.data_ptr(), m, n, k, 0, st))
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        err = None
        if want is not None:
            got = ty.cpu().numpy().astype(np.float64)
            err = (np.abs(got - want).max() / np.abs(want).max(), np.sqrt(((got - want) ** 2).mean()) / np.sqrt((want ** 2).mean()))
        res[name] = (us, err)
    ref = torch.matmul(tx, tw.t())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(tx, tw.t())
    e1.record()
    torch.cuda.synchronize()
    flops = 2.0 * m * n * k
    print("M=%6d N=%3d K=%3d | tc %7.1f us (%.1f TFLOP/s eff, x3 on the pipe) err %s | simt %7.1f us (%.1f TF/s) err %s | cuBLAS fp32 %7.1f us"
          % (m, n, k, res["tc"][0], flops / res["tc"][0] / 1e6, res["tc"][1], res["simt"][0], flops / res["simt"][0] / 1e6,
             res["simt"][1], e0.elapsed_time(e1) * 100), flush=True)
