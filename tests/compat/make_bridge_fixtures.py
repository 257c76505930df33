#!/usr/bin/env python
"""Engine inputs as the REFERENCE's own layer builds them -> tests/golden/bridge_<model>.npz.

Runs only where /root/reference exists: imports the reference's `gempy` (tests/compat/ref_harness.py), builds the example
models through its public API (gp.generate_example_model(..., compute_model=False), gempy/API/examples_generator.py) and
exports what its bridge (interpolation_input_from_structural_frame, _engine_factory.py:14-58), GeoModel.interpolation_options
and StructuralFrame.input_data_descriptor hand to the engine call (compute_API.py:68-73).  The GPU tests load these files
(tests/test_gpu_parity.py::test_reference_bridge_inputs_*), so the approved vectors are reproduced from inputs that went
through the reference's real code, not through this repo's restated example builders."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh                                   # noqa: E402

MODELS = ["HORIZONTAL_STRAT", "ANTICLINE", "ONE_FAULT", "COMBINATION", "FOLD", "RECUMBENT_FOLD", "PINCH_OUT", "GRABEN"]


def main():
    gp = rh.import_gempy()
    from gempy.core.data.enumerators import ExampleModel
    from gempy.modules.data_manipulation import interpolation_input_from_structural_frame
    from gempy_b200.engine.io import engine_inputs_to_npz
    out_dir = os.path.join(rh.ROOT, "tests", "golden")
    for name in MODELS:
        if not hasattr(ExampleModel, name):
            continue
        try:
            m = gp.generate_example_model(getattr(ExampleModel, name), compute_model=False)
        except Exception as exc:                           # models whose generators need more than the stand-ins offer
            print(f"{name}: skipped ({type(exc).__name__}: {exc})")
            continue
        m.validate()
        ii = interpolation_input_from_structural_frame(m)
        path = os.path.join(out_dir, f"bridge_{name.lower()}.npz")
        engine_inputs_to_npz(path, ii, m.interpolation_options, m.input_data_descriptor)
        print(f"{name}: {ii.surface_points.n_points} surface points, {ii.orientations.n_items} orientations -> {os.path.relpath(path, rh.ROOT)}")


if __name__ == "__main__":
    main()
