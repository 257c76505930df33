from ._types import (InterpolationOptions, Solutions, SurfacePoints, Orientations, InterpolationInput, TensorsStructure,
                     StacksStructure, InputDataDescriptor)      # noqa: F401
from . import engine_grid                                         # noqa: F401
