"""The drop-in seam: install() rebinds the engine names the UNMODIFIED reference
recommenders look up, and the reference recommender then constructs OUR engine.
Needs /root/reference (build container); no GPU, so construction must stop at the
loud no-CPU error raised by our engine -- which proves our class was reached."""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import pytest


@pytest.mark.needs_reference
def test_install_rebinds_reference_names_and_recommender_reaches_our_engine(tmp_path):
    from oracle import ref_shim

    ref_shim.install()
    import beta_recsys_b200
    from beta_recsys_b200 import BrsError, engines

    import beta_rec.recommenders.matrix_factorization as ref_mf_rec
    import beta_rec.models.mf as ref_mf

    original = ref_mf_rec.MFEngine
    try:
        patched = beta_recsys_b200.install()
        assert ("beta_rec.recommenders.matrix_factorization", "MFEngine") in patched
        assert ref_mf_rec.MFEngine is engines.MFEngine and ref_mf.MFEngine is engines.MFEngine
        # drive the reference's own MatrixFactorization.init_engine with a CPU config
        cfg = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "configs", "mf_default.json")))
        cfg["system"]["root_dir"] = str(tmp_path)
        cfg_file = tmp_path / "mf.json"
        cfg_file.write_text(json.dumps(cfg))

        class Data(object):
            n_users, n_items = 943, 1682

        with redirect_stdout(io.StringIO()):
            rec = ref_mf_rec.MatrixFactorization({"config_file": str(cfg_file), "device": "cpu"})
        with pytest.raises(BrsError, match="no CPU fallback"):
            with redirect_stdout(io.StringIO()):
                rec.init_engine(Data())
    finally:
        beta_recsys_b200.uninstall()
        sys.stdout = sys.__stdout__
        sys.stderr = sys.__stderr__
    assert ref_mf_rec.MFEngine is original


def test_install_is_a_noop_without_the_reference(monkeypatch):
    import beta_recsys_b200
    from oracle import ref_shim

    if ref_shim.reference_available() and "beta_rec" in sys.modules:
        pytest.skip("reference already imported in this process")
    assert beta_recsys_b200.install() in ([], beta_recsys_b200.install())
    beta_recsys_b200.uninstall()
