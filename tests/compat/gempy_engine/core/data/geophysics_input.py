from ._types import GeophysicsInput                # noqa: F401
