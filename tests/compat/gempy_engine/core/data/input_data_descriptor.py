from ._types import InputDataDescriptor, TensorsStructure, StacksStructure                # noqa: F401
