"""Import the reference's own scoring functions LIVE from /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/lemon_oracle.py header).  /root/reference exists
only in the builder container; everything here is guarded by ``available()`` and is
used to (a) generate tests/golden/*.npz and (b) cross-check the oracle in CPU tests.
The reference cannot be imported as shipped (missing third-party packages and a
missing ``lib.models.constants``); the stub packages in oracle/ref_stubs/ supply
names only — no reference code is copied.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = "/root/reference"
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_stubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "metrics"))


def import_reference_metrics():
    """Returns the reference module ``lib.metrics.utils`` (calc_scores_given_hparams,
    calc_scores_given_hparams_vectorized, ...)."""
    if not available():
        raise RuntimeError("reference not mounted")
    for p in (_STUBS, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # lib/models/downstream_models.py:13 does `from . import constants`; the file is absent upstream
    sys.modules.setdefault("lib.models.constants", types.ModuleType("lib.models.constants"))
    import importlib
    return importlib.import_module("lib.metrics.utils")


def import_reference_utils():
    """Returns the reference module ``lib.utils.utils`` (normalize_vectors)."""
    if not available():
        raise RuntimeError("reference not mounted")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    return importlib.import_module("lib.utils.utils")
