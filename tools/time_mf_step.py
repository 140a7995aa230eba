"""Time the MF step at BASELINE config 2 for both step implementations (CUDA events, device-resident
batches) and cross-check them against each other from identical state.

    python tools/time_mf_step.py [--impl rows scratch] [--chunk 8] [--opt sgd] [--users N --items N --dim D --batch B]
"""
import argparse
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


def make(a, impl):
    cfg = {"model": dict(device_str="cuda:0", n_users=a.users, n_items=a.items, emb_dim=a.dim, batch_size=a.batch,
                         optimizer=a.opt, lr=0.05, loss="bpr", adam_mode=a.adam_mode, step_impl=impl),
           "system": {"run_dir": "/tmp/brs_time"}}
    torch.manual_seed(2020)
    with redirect_stdout(io.StringIO()):
        return MFEngine(cfg)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", nargs="+", default=["rows", "scratch"])
    ap.add_argument("--shape", nargs="+", default=["-"], help="unused (kept for old command lines)")
    ap.add_argument("--opt", default="sgd")
    ap.add_argument("--adam-mode", default="touched")
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--items", type=int, default=100_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--nb", type=int, default=256)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--zipf", type=float, default=1.05)
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    bench.ZIPF_A = a.zipf
    dev = torch.device("cuda:0")
    lib = _lib.load()
    users, pos, neg = bench.make_batches(a.users, a.items, a.batch, a.nb, 2020, dev)
    stream = torch.cuda.current_stream(dev)
    n = a.nb * a.batch

    if not a.no_check and "rows" in a.impl and "scratch" in a.impl:
        e1, e2 = make(a, "rows"), make(a, "scratch")
        with torch.no_grad():
            for (k, v), (_, w) in zip(e1.model.state_dict().items(), e2.model.state_dict().items()):
                w.copy_(v)
        s0 = {k: v.clone() for k, v in e1.model.state_dict().items()}
        m = 4 * a.batch
        r1 = e1.train_batches(users[:m], pos[:m], neg[:m])
        r2 = e2.train_batches(users[:m], pos[:m], neg[:m])
        print("records rows   :", r1[:, :2].tolist())
        print("records scratch:", r2[:, :2].tolist())
        for (k, v), (_, w) in zip(e1.model.state_dict().items(), e2.model.state_dict().items()):
            dw = (v - s0[k]).abs().max().item()
            print("  %-18s max|dw| %.3e  max|rows-scratch| %.3e  rel-to-dw %.2e" %
                  (k, dw, (v - w).abs().max().item(), (v - w).abs().max().item() / max(dw, 1e-30)))
        del e1, e2, s0
        torch.cuda.empty_cache()

    for impl in a.impl:
        for shape in ["-"]:
            eng = make(a, impl)
            out = torch.empty((a.nb, 4), dtype=torch.float32, device=dev)

            def run():
                _lib.check(lib.brs_mf_train_batches(eng._cmodel, eng.optimizer.desc, 0, _lib.ptr(users), _lib.ptr(pos),
                                                    _lib.ptr(neg), n, a.batch, 0.0, _lib.ptr(out), stream.cuda_stream))

            run()
            torch.cuda.synchronize()
            t = []
            for _ in range(a.passes):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                run()
                e1.record(stream)
                torch.cuda.synchronize()
                t.append(e0.elapsed_time(e1) / a.nb * 1e3)
            us = float(np.median(t))
            st = _lib.step_records_status(out.cpu().numpy())
            alg = (24 * a.dim + 48) * a.batch
            print("impl %-7s shape %-8s  %s  step %.1f us  %.0f M inter/s  alg %.0f GB/s  status %d  loss %.5f" %
                  (impl, shape, a.opt, us, a.batch / us, alg / us / 1e3, st, out[-1, 0].item()))
            if impl == "rows":  # serial phases of one step
                k = 40
                ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(k)]
                o1 = torch.empty(4, dtype=torch.float32, device=dev)
                for q in range(k):
                    off = (q % a.nb) * a.batch
                    ev[q][0].record(stream)
                    _lib.check(lib.brs_mf_plan_build(eng._cmodel, 0, 0, _lib.ptr(users[off:]), _lib.ptr(pos[off:]),
                                                     _lib.ptr(neg[off:]), a.batch, stream.cuda_stream))
                    ev[q][1].record(stream)
                    _lib.check(lib.brs_mf_step_planned(eng._cmodel, 0, eng.optimizer.desc, 0, a.batch, 0.0, _lib.ptr(o1),
                                                       stream.cuda_stream))
                    ev[q][2].record(stream)
                torch.cuda.synchronize()
                tp = np.median([e[0].elapsed_time(e[1]) for e in ev]) * 1e3
                ts = np.median([e[1].elapsed_time(e[2]) for e in ev]) * 1e3
                print("      serial: plan (claim+scan+fill) %.1f us, rows (users+items) %.1f us" % (tp, ts))
                # each row kernel alone, repeated on one fixed plan (weights drift; timing only)
                _lib.check(lib.brs_mf_plan_build(eng._cmodel, 0, 0, _lib.ptr(users), _lib.ptr(pos), _lib.ptr(neg), a.batch,
                                                 stream.cuda_stream))
                for which, name in ((1, "users"), (2, "items")):
                    lib.brs_debug_set_mf_rows_only(which)
                    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
                    ev2[0].record(stream)
                    for q in range(20):
                        _lib.check(lib.brs_mf_step_planned(eng._cmodel, 0, eng.optimizer.desc, 0, a.batch, 0.0, _lib.ptr(o1),
                                                           stream.cuda_stream))
                        ev2[q + 1].record(stream)
                    torch.cuda.synchronize()
                    print("      %s kernel alone (same plan, warm L2): median %.1f us" %
                          (name, np.median([ev2[q].elapsed_time(ev2[q + 1]) for q in range(20)]) * 1e3))
                lib.brs_debug_set_mf_rows_only(0)
            del eng
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
