from ._types import InterpolationInput                # noqa: F401
