"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a without a
GPU, loads, and exports every symbol include/brs_b200.h declares; the ctypes
struct mirrors have the C layout; the product refuses to run without CUDA."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "brs_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    from beta_recsys_b200 import build

    return build.build()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(brs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_survey_minimum_set():
    names = declared_functions()
    for need in ("brs_mf_bpr_fwd_bwd", "brs_mf_bce_fwd_bwd", "brs_rows_sgd", "brs_rows_adam", "brs_dense_adam_sweep",
                 "brs_gather", "brs_scatter_add"):
        assert need in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), "missing export: " + name
    assert lib.brs_abi_version() == 2
    lib.brs_strerror.restype = ctypes.c_char_p
    assert lib.brs_strerror(0) == b"ok"
    assert b"unsupported" in lib.brs_strerror(-2)


def test_ctypes_prototypes_cover_the_header(lib_path):
    from beta_recsys_b200 import _lib

    assert sorted(_lib._PROTOTYPES) == declared_functions()
    _lib.load()


def test_struct_layout_matches_c(lib_path, tmp_path):
    """sizeof/offsetof of every ABI struct, as gcc sees the header, equals the ctypes mirror."""
    from beta_recsys_b200 import _lib

    structs = {"brs_opt": _lib.Opt, "brs_table": _lib.Table, "brs_rowset": _lib.Rowset, "brs_entity": _lib.Entity,
               "brs_dense_param": _lib.DenseParam, "brs_mf_model": _lib.MfModel}
    extra = getattr(_lib, "EXTRA_STRUCTS", {})
    structs.update(extra)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "brs_b200.h"', "int main(void){"]
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(out[cname + "." + f]) == getattr(cls, f).offset, (cname, f)


def test_no_device_is_reported_not_crashed(lib_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from beta_recsys_b200 import _lib

    lib = _lib.load()
    sm = ctypes.c_int(0)
    assert lib.brs_device_info(0, ctypes.byref(sm), None, None, None) == -5  # BRS_ERR_NO_DEVICE


def test_engine_refuses_cpu_loudly():
    from beta_recsys_b200 import BrsError
    from beta_recsys_b200.engines import MFEngine

    cfg = {"model": dict(device_str="cpu", n_users=5, n_items=5, emb_dim=8, batch_size=4, optimizer="sgd", lr=0.1),
           "system": {"run_dir": "/tmp/x"}}
    with pytest.raises(BrsError, match="no CPU fallback"):
        MFEngine(cfg)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under beta_recsys_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "beta_recsys_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "cf_oracle" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_sass_has_tma_and_vector_red(lib_path):
    """The staged-index path really is a TMA bulk copy and the scatter is a 128-bit RED."""
    try:
        sass = subprocess.check_output(["cuobjdump", "-sass", lib_path], text=True)
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    assert "UBLKCP" in sass
    assert "REDG.E.ADD.F32x4" in sass
    assert "LDG.E.NA.128" in sass or "LDG.E.128" in sass
