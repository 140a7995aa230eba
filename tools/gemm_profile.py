#!/usr/bin/env python
"""One shape of the tensor-core Linear kernel, a few launches: the target of an ncu capture.

    ncu --set full --import-source on -k regex:linear_tc --launch-skip 2 --launch-count 1 -o prof python tools/gemm_profile.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from beta_recsys_b200 import _lib  # noqa: E402

m, n, k = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (65536, 256, 512)))
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda")
g.manual_seed(0)
x = torch.randn((m, k), device="cuda", generator=g)
w = torch.randn((n, k), device="cuda", generator=g) / k ** 0.5
b = torch.zeros(n, device="cuda")
y = torch.empty((m, n), device="cuda")
for _ in range(4):
    _lib.check(lib.brs_mlp_fwd_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), None, m, n, k, 0, st))
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
