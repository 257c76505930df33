"""Drive the reference's own `gempy` layer (from /root/reference, read-only) against the B200 backend: the `gempy_engine`
stand-in of this directory on sys.path (plus a `pooch` stand-in that resolves the example generators' data URLs to the
reference's local copies of the CSVs), and the backend arm of INTEGRATION.md installed (gempy_b200.integration.install_backend_arm).

Only usable where /root/reference exists (this container, no GPU); the GPU box gets the engine inputs this harness
exports (tests/golden/bridge_*.npz, written by tests/compat/make_bridge_fixtures.py)."""
import os
import sys

REFERENCE = os.environ.get("GEMPY_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "gempy"))


def import_gempy():
    """`import gempy` from the reference tree with the gempy_engine stand-in; returns the module."""
    for p in (ROOT, HERE, REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pandas as pd
    try:                                     # the reference pins pandas < 3 (requirements/base-requirements.txt): with
        pd.set_option("future.infer_string", False)      # pandas 3 string columns must come back as object arrays
    except Exception:
        pass
    import gempy
    from gempy_b200 import integration
    integration.install_backend_arm(gempy)
    return gempy
