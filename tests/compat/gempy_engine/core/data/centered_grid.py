from ._types import CenteredGrid                # noqa: F401
