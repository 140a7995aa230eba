"""GPU-bound timing of the row-owner MF kernels: each call sequence is captured into a CUDA graph and
replayed, so that host launch overhead (Python, ctypes, driver) is out of the measurement.

    python tools/exp_rows.py
"""
import io
import os
import sys
from contextlib import redirect_stdout

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from beta_recsys_b200 import _lib  # noqa: E402
from beta_recsys_b200.engines import MFEngine  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
B = 65536


def engine(impl):
    cfg = {"model": dict(device_str="cuda:0", n_users=1_000_000, n_items=100_000, emb_dim=128, batch_size=B, optimizer="sgd",
                         lr=0.05, loss="bpr", step_impl=impl), "system": {"run_dir": "/tmp/x"}}
    with redirect_stdout(io.StringIO()):
        return MFEngine(cfg)


users, pos, neg = bench.make_batches(1_000_000, 100_000, B, 8, 2020, dev)
o1 = torch.zeros(4, device=dev)


def timed_graph(fn, reps=50, inner=1):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(s.cuda_stream)  # warm-up outside capture (attribute calls, lazy init)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(inner):
            fn(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * inner) * 1e3


eng = engine("rows")
st0 = torch.cuda.current_stream().cuda_stream
_lib.check(lib.brs_mf_plan_build(eng._cmodel, 0, 0, _lib.ptr(users), _lib.ptr(pos), _lib.ptr(neg), B, st0))
torch.cuda.synchronize()
for which, name in ((1, "users"), (2, "items")):
    lib.brs_debug_set_mf_rows_only(which)
    us = timed_graph(lambda st: _lib.check(lib.brs_mf_step_planned(eng._cmodel, 0, eng.optimizer.desc, 0, B, 0.0, _lib.ptr(o1), st)),
                     inner=4)
    print("rows %-5s kernel alone, graph replay (fixed plan, warm L2): %.1f us" % (name, us))
lib.brs_debug_set_mf_rows_only(0)
us = timed_graph(lambda st: _lib.check(lib.brs_mf_plan_build(eng._cmodel, 0, 0, _lib.ptr(users), _lib.ptr(pos), _lib.ptr(neg), B, st)), inner=1)
print("rows plan_build alone (graph replay): %.1f us  (NB: replays re-claim already claimed slots)" % us)


def full(st):
    for k in range(8):
        o = k * B
        _lib.check(lib.brs_mf_step(eng._cmodel, eng.optimizer.desc, 0, _lib.ptr(users[o:]), _lib.ptr(pos[o:]), _lib.ptr(neg[o:]), B, 0.0,
                                   _lib.ptr(o1), st))


torch.cuda.synchronize()
us = timed_graph(full, reps=20, inner=1) / 8
print("rows full step (plan + users + items serial on one stream, 8 distinct batches per graph): %.1f us/step" % us)
del eng
eng2 = engine("scratch")


def full2(st):
    for k in range(8):
        o = k * B
        _lib.check(lib.brs_mf_bpr_fwd_bwd(eng2._cmodel, _lib.ptr(users[o:]), _lib.ptr(pos[o:]), _lib.ptr(neg[o:]), B, 0.0, st))
        _lib.check(lib.brs_mf_apply(eng2._cmodel, eng2.optimizer.desc, B, _lib.ptr(o1), st))


us = timed_graph(full2, reps=20, inner=1) / 8
print("scratch full step (prepass + fwd_bwd + apply serial, 8 distinct batches per graph): %.1f us/step" % us)
