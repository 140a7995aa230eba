import sys, argparse
sys.path.insert(0, "/root/repo")
import torch, bench
a = argparse.Namespace(users=10_000_000, items=1_000_000, dim=128, batch=65536, optimizer="sgd", adam_mode="dense")
print(bench.single_gpu_same_tables(a, torch.device("cuda", 0)))
