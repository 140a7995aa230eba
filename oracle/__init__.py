"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the beta_rec embedding-CF training step.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline.  The product path (``beta_recsys_b200``) never imports this package
and fails loudly when its CUDA library is missing.

Parity pin: the reference's own tests never touch a model, loss or optimizer
 so there are no reference-held golden vectors.  The oracle is
pinned instead against outputs of the reference itself, run in the build
container through ``oracle/ref_shim.py`` by ``oracle/make_golden.py``; those
outputs are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py`` on every run.
"""
