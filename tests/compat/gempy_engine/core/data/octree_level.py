from ._types import OctreeLevel                # noqa: F401
