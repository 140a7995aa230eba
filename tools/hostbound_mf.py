import io, sys, time
from contextlib import redirect_stdout
sys.path.insert(0, "/root/repo")
import torch, bench
from beta_recsys_b200 import _lib
from beta_recsys_b200.engines import MFEngine
lib = _lib.load(); dev = torch.device("cuda:0"); B = 65536; nb = 256
cfg = {"model": dict(device_str="cuda:0", n_users=1_000_000, n_items=100_000, emb_dim=128, batch_size=B, optimizer="sgd", lr=0.05, loss="bpr"), "system": {"run_dir": "/tmp/x"}}
with redirect_stdout(io.StringIO()):
    eng = MFEngine(cfg)
users, pos, neg = bench.make_batches(1_000_000, 100_000, B, nb, 2020, dev)
out = torch.empty((nb, 4), device=dev)
st = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(lib.brs_mf_train_batches(eng._cmodel, eng.optimizer.desc, 0, _lib.ptr(users), _lib.ptr(pos), _lib.ptr(neg), nb * B, B, 0.0, _lib.ptr(out), st))
run(); torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); run(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("host enqueue %.1f us/step, total %.1f us/step" % ((t1 - t0) / nb * 1e6, (t2 - t0) / nb * 1e6))
