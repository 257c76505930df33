from ._types import Transform, GlobalAnisotropy                # noqa: F401
