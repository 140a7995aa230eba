"""Per-warp phase cycle counters of the MF users kernel (build with BRS_NVCC_DEFINES=-DBRS_ROWS_PROFILE).

    BRS_NVCC_DEFINES=-DBRS_ROWS_PROFILE python -m beta_recsys_b200.build --force && python tools/prof_mf_rows.py
"""
import io
import os
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from beta_recsys_b200 import _lib  # noqa: E402
from beta_recsys_b200.engines import MFEngine  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
B = 65536
cfg = {"model": dict(device_str="cuda:0", n_users=1_000_000, n_items=100_000, emb_dim=128, batch_size=B, optimizer="sgd",
                     lr=0.05, loss="bpr"), "system": {"run_dir": "/tmp/x"}}
with redirect_stdout(io.StringIO()):
    eng = MFEngine(cfg)
users, pos, neg = bench.make_batches(1_000_000, 100_000, B, 8, 2020, dev)
_lib.check(lib.brs_debug_mf_rows_profile(None, 0))
if len(sys.argv) > 1:
    _lib.check(lib.brs_debug_set_mf_rows_shape(*[int(x) for x in sys.argv[1].split(",")]))
out = torch.zeros(4, device=dev)
st = torch.cuda.current_stream().cuda_stream
for k in range(4):
    o = k * B
    _lib.check(lib.brs_mf_step(eng._cmodel, eng.optimizer.desc, 0, _lib.ptr(users[o:]), _lib.ptr(pos[o:]), _lib.ptr(neg[o:]), B,
                               0.0, _lib.ptr(out), st))
torch.cuda.synchronize()
n = 148 * 16
buf = np.zeros((n, 8), dtype=np.int64)
_lib.check(lib.brs_debug_mf_rows_profile(buf.ctypes.data, n))
buf = buf[buf[:, 0] > 0]
names = ["total", "tile rec load", "issue", "wait+sync", "lds+dots+shfl", "chain", "flush", "iterations"]
print("warps with work:", len(buf))
for k, nm in enumerate(names):
    c = buf[:, k]
    print("%-16s mean %9.0f  p50 %9.0f  p90 %9.0f  max %9.0f" % (nm, c.mean(), np.median(c), np.percentile(c, 90), c.max()))
it = buf[:, 7].clip(min=1)
print("per iteration: total %.0f cycles; issue %.0f wait %.0f dots %.0f chain %.0f; rest %.0f" % (
    (buf[:, 0] / it).mean(), (buf[:, 2] / it).mean(), (buf[:, 3] / it).mean(), (buf[:, 4] / it).mean(), (buf[:, 5] / it).mean(),
    ((buf[:, 0] - buf[:, 1] - buf[:, 2] - buf[:, 3] - buf[:, 4] - buf[:, 5] - buf[:, 6]) / it).mean()))
