from ._types import InterpOutput                # noqa: F401
