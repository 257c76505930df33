from ._types import StackRelationType                # noqa: F401
