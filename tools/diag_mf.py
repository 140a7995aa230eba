#!/usr/bin/env python
"""Round-2 diagnostic battery for the single-GPU MF step (one gpurun call, ~2 min): how the three kernels
of the step react to the popularity law of the stream and to the table size.  Answers, with measurements:

  * what does the heavy head of Zipf(1.05) cost (one user = 9.5 % of a batch, profiles/r01_batch_stats.md)?
    -> compare --zipf-a 1.05 / 0.8 / 0 (uniform) at the benchmark shape;
  * how far is the fused kernel from its all-in-L2 floor?  -> small tables (100k x 10k);
  * does D change the picture?  -> D = 64 / 128 / 256.

    python tools/diag_mf.py > gpurun_out/diag_mf.md
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "400", "--warmup", "10", "--no-e2e",
           "--no-cpu-baseline"] + extra
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if out.returncode != 0 or not lines:
        return None, (out.stderr or out.stdout)[-300:]
    return json.loads(lines[-1]), None


def main():
    cases = [
        ("benchmark shape, Zipf 1.05", []),
        ("benchmark shape, Zipf 0.8", ["--zipf-a", "0.8"]),
        ("benchmark shape, uniform ids", ["--zipf-a", "0"]),
        ("100k x 10k tables (L2-resident), Zipf 1.05", ["--users", "100000", "--items", "10000"]),
        ("100k x 10k tables (L2-resident), uniform", ["--users", "100000", "--items", "10000", "--zipf-a", "0"]),
        ("D = 64", ["--dim", "64"]),
        ("D = 256", ["--dim", "256"]),
        ("Adam, touched rows", ["--optimizer", "adam", "--adam-mode", "touched"]),
    ]
    print("| case | M inter/s | us/step | pre-pass us | fused us | apply us | fused GB/s (algorithmic) | frac |")
    print("|---|---|---|---|---|---|---|---|")
    for name, extra in cases:
        d, err = run(extra)
        if d is None:
            print("| %s | failed: %s |" % (name, err.replace("\n", " ")[:120]))
            continue
        r = d["roofline"]
        print("| %s | %.0f | %.1f | %.1f | %.1f | %.1f | %.0f | %.2f |"
              % (name, d["value"] / 1e6, d["ms_per_step"] * 1e3, r["prepass_kernel_ms"] * 1e3, r["kernel_ms"] * 1e3,
                 r["apply_kernel_ms"] * 1e3, r["achieved"], r["frac"]), flush=True)


if __name__ == "__main__":
    main()
