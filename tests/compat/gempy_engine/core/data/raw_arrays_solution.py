from ._types import RawArraysSolution                # noqa: F401
