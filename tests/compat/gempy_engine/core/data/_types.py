"""gempy_b200's data model under the names `gempy` imports, each made acceptable as a pydantic field type (GeoModel,
StructuralFrame and Grid are pydantic models whose fields are annotated with engine classes,
gempy/core/data/geo_model.py:56-77): instances are passed through unchanged."""
from pydantic_core import core_schema

from gempy_b200.engine import data as _d
from gempy_b200.engine import geophysics as _g
_d.GeophysicsInput = _g.GeophysicsInput


def _passthrough(cls):
    def __get_pydantic_core_schema__(klass, source, handler):
        return core_schema.is_instance_schema(klass, serialization=core_schema.plain_serializer_function_ser_schema(
            lambda v: getattr(v, "__dict__", str(v)), when_used="json"))
    cls.__get_pydantic_core_schema__ = classmethod(__get_pydantic_core_schema__)
    return cls


for _name in ("Transform", "InterpolationOptions", "KernelOptions", "EvaluationOptions", "Solutions", "RawArraysSolution",
              "EngineGrid", "RegularGrid", "GenericGrid", "CenteredGrid", "InterpolationInput", "InputDataDescriptor",
              "TensorsStructure", "StacksStructure", "SurfacePoints", "Orientations", "FaultsData", "OctreeLevel",
              "InterpOutput", "DualContouringMesh", "GeophysicsInput"):
    if hasattr(_d, _name):
        globals()[_name] = _passthrough(getattr(_d, _name))

StackRelationType = _d.StackRelationType
AvailableKernelFunctions = _d.AvailableKernelFunctions
BlockSolutionType = _d.BlockSolutionType
GlobalAnisotropy = _d.GlobalAnisotropy
FiniteFaultData = getattr(_d, "FiniteFaultData", None)
