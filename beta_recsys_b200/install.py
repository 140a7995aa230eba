"""Drop the B200 engines in under the UNMODIFIED reference recommenders.

The reference has no plugin registry: each recommender imports its engine class by
name and looks the name up at call time (e.g. ``MFEngine`` at
beta_rec/recommenders/matrix_factorization.py:8,74).  ``install()`` rebinds those
module-level names -- no reference file is edited -- so that
``MatrixFactorization(config).train(data)``, ``NeuCF(config).train(data)`` and the
example scripts construct the CUDA engines instead.  ``uninstall()`` restores them.
"""
import importlib

# (module that holds the name, attribute, replacement factory)
_BINDINGS = [
    ("beta_rec.recommenders.matrix_factorization", "MFEngine", "MFEngine"),
    ("beta_rec.models.mf", "MFEngine", "MFEngine"),
    ("beta_rec.recommenders.ncf", "NeuMFEngine", "NeuMFEngine"),
    ("beta_rec.models.ncf", "NeuMFEngine", "NeuMFEngine"),
    ("beta_rec.models.gmf", "GMFEngine", "GMFEngine"),
    ("beta_rec.models.mlp", "MLPEngine", "MLPEngine"),
    ("beta_rec.recommenders.lightgcn", "LightGCNEngine", "LightGCNEngine"),
    ("beta_rec.models.lightgcn", "LightGCNEngine", "LightGCNEngine"),
    # ranking evaluation: train_eval_worker / test_eval_worker look `evaluate` up at call time
    # (beta_rec/core/eval_engine.py:106,109)
    ("beta_rec.core.eval_engine", "evaluate", "eval.evaluate"),
]
_saved = {}


def install(strict=False):
    """Rebind the reference's engine names to the B200 engines.  Returns the list of
    (module, attribute) pairs that were patched.  Modules that cannot be imported
    (beta_rec not installed, missing optional deps) are skipped unless ``strict``."""
    from . import engines

    done = []
    for mod_name, attr, ours in _BINDINGS:
        try:
            mod = importlib.import_module(mod_name)
        except Exception:
            if strict:
                raise
            continue
        if not hasattr(mod, attr):
            if strict:
                raise AttributeError("%s has no attribute %s" % (mod_name, attr))
            continue
        _saved.setdefault((mod_name, attr), getattr(mod, attr))
        if ours.startswith("eval."):
            from . import eval as _eval

            setattr(mod, attr, getattr(_eval, ours[5:]))
        else:
            setattr(mod, attr, getattr(engines, ours))
        done.append((mod_name, attr))
    return done


def uninstall():
    for (mod_name, attr), orig in list(_saved.items()):
        setattr(importlib.import_module(mod_name), attr, orig)
        del _saved[(mod_name, attr)]
