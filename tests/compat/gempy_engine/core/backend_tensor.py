"""gempy_engine.core.backend_tensor: the process-global backend switch (gempy/API/compute_API.py:45-50,
gempy/modules/optimize_nuggets/_optimizer.py:24).  The B200 backend has one numeric mode (float64 on the device), so the
switch records what was asked and does nothing else."""
import numpy as np
from ..config import AvailableBackends


class BackendTensor:
    engine_backend = AvailableBackends.numpy
    use_gpu = False
    dtype = "float64"
    dtype_obj = np.float64
    tfnp = np
    t = np
    PYKEOPS = False
    COMPUTE_GRADS = False

    @classmethod
    def change_backend_gempy(cls, engine_backend, use_gpu=False, dtype=None, grads=False):
        cls.engine_backend, cls.use_gpu, cls.COMPUTE_GRADS = engine_backend, bool(use_gpu), bool(grads)
        if dtype is not None:
            cls.dtype = dtype
