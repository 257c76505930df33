"""gempy_engine.config: backend selector (gempy/core/data/gempy_engine_config.py:5-11, gempy/API/compute_API.py:41-42)."""
import os
from gempy_b200.engine.data import AvailableBackends              # noqa: F401  numpy / PYTORCH / legacy + B200

DEFAULT_BACKEND = AvailableBackends[os.getenv("DEFAULT_BACKEND", "numpy")] if os.getenv("DEFAULT_BACKEND", "numpy") in AvailableBackends.__members__ else AvailableBackends.numpy
DEFAULT_PYKEOPS = False
DEFAULT_TENSOR_DTYPE = "float64"
