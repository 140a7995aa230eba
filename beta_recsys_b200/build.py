"""Build libbrs_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m beta_recsys_b200.build [--force] [--verbose]

The shared library has a plain C ABI (include/brs_b200.h), links the CUDA runtime
statically and has no dependency on torch.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbrs_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libbrs_b200.so")


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))


def build(force=False, verbose=False):
    """One object per .cu (compiled in parallel, rebuilt only when the source or a header is newer),
    then one link.  BRS_NVCC_DEFINES passes extra -D flags (experiments only)."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    nvcc = find_nvcc()
    extra = os.environ.get("BRS_NVCC_DEFINES", "").split()
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, ".defines")
    if not os.path.exists(stamp) or open(stamp).read() != " ".join(extra):
        force = True
    hdr_t = max([os.path.getmtime(p) for p in _headers()] + [os.path.getmtime(os.path.abspath(__file__))])

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, 0, ""
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return obj, r.returncode, ("$ " + " ".join(cmd) + "\n" + r.stdout)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    failed = [r for r in results if r[1] != 0]
    for _, rc, log in results:
        if log and (verbose or rc != 0):
            print(log)
    if failed:
        raise RuntimeError("nvcc failed building libbrs_b200.so")
    cmd = [nvcc] + LINK_FLAGS + ["-o", LIB + ".tmp"] + [r[0] for r in results]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed linking libbrs_b200.so")
    os.replace(LIB + ".tmp", LIB)
    with open(stamp, "w") as f:
        f.write(" ".join(extra))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
