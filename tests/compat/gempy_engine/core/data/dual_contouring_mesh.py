from ._types import DualContouringMesh                # noqa: F401
