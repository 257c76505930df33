from ._types import Solutions                # noqa: F401
