#!/bin/bash
# Put the UNMODIFIED reference package where the GPU box can import it: baseline/_ref (git-ignored, travels with
# gpurun).  Used only by tests marked needs_reference (tests/test_reference_train_gpu.py: the reference's own
# MatrixFactorization.train() running on the B200 engines through install()).
#   1. the contract's pip install (fails here: setup.py's setup_requires=pytest-runner is not in the wheelhouse)
#   2. fallback: copy the package directory, which is all `pip install --target` would have placed there
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${BETA_REC_REFERENCE:-/root/reference}"
[ -d "$REF/beta_rec" ] || { echo "no reference at $REF"; exit 0; }
rm -rf "$ROOT/baseline/_ref" /tmp/brs_refcopy
mkdir -p "$ROOT/baseline"
cp -r "$REF" /tmp/brs_refcopy
if python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target "$ROOT/baseline/_ref" /tmp/brs_refcopy >/tmp/brs_ref_pip.log 2>&1; then
    echo "pip install ok"
else
    echo "pip install failed (see /tmp/brs_ref_pip.log); copying the package directory instead"
    mkdir -p "$ROOT/baseline/_ref"
    cp -r "$REF/beta_rec" "$ROOT/baseline/_ref/"
fi
rm -rf /tmp/brs_refcopy
