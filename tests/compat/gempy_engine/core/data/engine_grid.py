from ._types import EngineGrid, RegularGrid, GenericGrid, CenteredGrid   # noqa: F401
