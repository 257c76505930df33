"""The reference-side binding of the B200 backend, as code.

INTEGRATION.md shows the `case AvailableBackends.B200:` arm a GemPy maintainer adds beside
/root/reference/gempy/API/compute_API.py:42.  `install_backend_arm(gempy)` applies exactly that arm to an imported
`gempy` without editing its source: `gp.compute_model(model, GemPyEngineConfig(backend=AvailableBackends.B200))` then runs

    gempy_model.validate()                                              (compute_API.py:38-39)
    interpolation_input_from_structural_frame(gempy_model)              (compute_API.py:65, _engine_factory.py:14-58)
    gempy_b200.engine.compute.compute_model(interpolation_input, options, data_descriptor, geophysics_input)
    gempy_model.solutions = <Solutions>                                 (geo_model.py:100-127)

and every other backend value falls through to the reference's own function."""
from __future__ import annotations

import functools

from .engine.compute import compute_model as _b200_compute_model
from .engine.data import AvailableBackends


def compute_model_b200(gempy_model, engine_config=None, skip_validation: bool = False, *, device=None, **kwargs):
    """The body of the new `case` arm (mirrors compute_API.py:36-39, 65-73 for this backend)."""
    from gempy.modules.data_manipulation import interpolation_input_from_structural_frame
    if not skip_validation:
        gempy_model.validate()
    interpolation_input = interpolation_input_from_structural_frame(gempy_model)
    gempy_model.taped_interpolation_input = interpolation_input
    gempy_model.solutions = _b200_compute_model(
        interpolation_input=interpolation_input,
        options=gempy_model.interpolation_options,
        data_descriptor=gempy_model.input_data_descriptor,
        geophysics_input=gempy_model.geophysics_input,
        device=device,
    )
    return gempy_model.solutions


def install_backend_arm(gempy_module) -> None:
    """Wrap gempy.API.compute_API.compute_model (and its re-exports gempy.compute_model / gempy.API.compute_model) with
    the B200 arm.  Idempotent."""
    api = gempy_module.API.compute_API
    original = api.compute_model
    if getattr(original, "_gpb_arm", False):
        return

    @functools.wraps(original)
    def compute_model(gempy_model, engine_config=None, skip_validation: bool = False, **kwargs):
        backend = getattr(engine_config, "backend", None)
        if getattr(backend, "name", None) == AvailableBackends.B200.name:
            return compute_model_b200(gempy_model, engine_config, skip_validation, **kwargs)
        return original(gempy_model, engine_config, skip_validation, **kwargs)

    compute_model._gpb_arm = True
    api.compute_model = compute_model
    for mod in (gempy_module, gempy_module.API):
        if getattr(mod, "compute_model", None) is original:
            mod.compute_model = compute_model
