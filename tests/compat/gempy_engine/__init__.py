"""TEST-ONLY stand-in for the `gempy_engine` package so that the reference's own `gempy` layer (bridge, GeoModel,
compute_API dispatch) can be imported and driven against the B200 backend.

The real engine (gempy_engine>=2026.0.3, /root/reference/requirements/requirements.txt:2) is not in the reference tree and
cannot be installed here.  `gempy` imports 22 distinct paths from it (grep in /root/reference/gempy); every one of them is
re-exported below from gempy_b200's own data model -- nothing is restated here except glue pydantic needs.  Used by
tests/compat/ only; the product never imports this package."""
from . import config                                              # noqa: F401
from gempy_b200.engine.compute import compute_model               # noqa: F401  (the engine entry point GemPy calls)
