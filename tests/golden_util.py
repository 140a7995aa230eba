"""Loader for tests/golden/*.npz (written by oracle/make_golden.py)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix=""):
    return sorted(
        os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz"))
    )


class Golden(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.name = name
        self._g = {}
        for k in z.files:
            if k == "meta":
                continue
            grp, rest = k.split("/", 1)
            self._g.setdefault(grp, {})[rest] = z[k]

    def group(self, g):
        return {k: v.copy() for k, v in self._g.get(g, {}).items()}

    @property
    def init(self):
        return self.group("init")

    @property
    def batch(self):
        return self.group("batch")

    @property
    def out(self):
        return self.group("out")


def max_rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): relative to the tensor's scale, the
    reading of '1e-5 relative' used throughout (elementwise relative error is
    meaningless for entries that are ~0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / scale)
