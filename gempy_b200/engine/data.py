"""Host-side data model of the B200 backend.

These classes mirror, attribute for attribute, the objects GemPy's bridge builds for
``gempy_engine.compute_model`` and the objects GemPy reads back from the returned
``Solutions`` (reference call site: gempy/API/compute_API.py:68-73; builders:
gempy/modules/data_manipulation/_engine_factory.py:26-56,71-104; consumers:
gempy/core/data/geo_model.py:100-127, gempy/modules/mesh_extranction/marching_cubes.py:27-47).
The engine package itself is not vendored in the reference tree, so the names below are the
ones the reference *uses*, not a copy of any source file.

Everything here is plain numpy on the host.  Device memory is owned by torch tensors inside
``gempy_b200.engine.compute`` and only materialised into these containers at the end of a
compute call.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


# --------------------------------------------------------------------------- enums
class AvailableBackends(enum.Enum):
    """Backend selector carried by ``GemPyEngineConfig.backend``
    (gempy/core/data/gempy_engine_config.py:11).  ``numpy`` / ``PYTORCH`` / ``legacy`` are the
    reference's members (gempy/API/compute_API.py:42, test/test_api/test_backends.py:55);
    ``B200`` is the member this backend adds."""
    numpy = enum.auto()
    PYTORCH = enum.auto()
    legacy = enum.auto()
    B200 = enum.auto()


class AvailableKernelFunctions(enum.Enum):
    cubic = 0
    exponential = 1
    matern_5_2 = 2


class StackRelationType(enum.Enum):
    """Encodings follow the serialization goldens (ERODE=1, FAULT=3;
    test/test_modules/test_faults/*.verify/fault.approved.txt)."""
    ERODE = 1
    ONLAP = 2
    FAULT = 3
    BASEMENT = 4


class BlockSolutionType(enum.Enum):
    """RawArraysSolution.BlockSolutionType (gempy/core/data/geo_model.py:324-339)."""
    NONE = 0
    OCTREE = 1
    DENSE_GRID = 2


class MeshExtractionMaskingOptions(enum.Enum):
    NOTHING = 1
    DISJOINT = 2
    INTERSECT = 3
    RAW = 4


# --------------------------------------------------------------------------- lazy host arrays
class Deferred:
    """A device-resident result that is copied to the host on first use.  ``compute_model`` returns in
    milliseconds at octree level 8; the multi-GB arrays only cross PCIe if somebody reads them."""
    __slots__ = ("_fetch", "_value")

    def __init__(self, fetch):
        self._fetch = fetch
        self._value = None

    def get(self):
        if self._fetch is not None:
            self._value = self._fetch()
            self._fetch = None
        return self._value


def _res(x):
    return x.get() if isinstance(x, Deferred) else x


# --------------------------------------------------------------------------- inputs
@dataclass
class SurfacePoints:
    """_engine_factory.py:27-30."""
    sp_coords: np.ndarray
    nugget_effect_scalar: np.ndarray | float = 2e-5

    def __post_init__(self):
        self.sp_coords = np.ascontiguousarray(self.sp_coords, dtype=np.float64).reshape(-1, 3)
        n = self.sp_coords.shape[0]
        self.nugget_effect_scalar = np.broadcast_to(
            np.asarray(self.nugget_effect_scalar, dtype=np.float64), (n,)).copy()

    @property
    def n_points(self) -> int:
        return self.sp_coords.shape[0]


@dataclass
class Orientations:
    """_engine_factory.py:33-37."""
    dip_positions: np.ndarray
    dip_gradients: np.ndarray
    nugget_effect_grad: np.ndarray | float = 0.01

    def __post_init__(self):
        self.dip_positions = np.ascontiguousarray(self.dip_positions, dtype=np.float64).reshape(-1, 3)
        self.dip_gradients = np.ascontiguousarray(self.dip_gradients, dtype=np.float64).reshape(-1, 3)
        n = self.dip_positions.shape[0]
        self.nugget_effect_grad = np.broadcast_to(
            np.asarray(self.nugget_effect_grad, dtype=np.float64), (n,)).copy()

    @property
    def n_items(self) -> int:
        return self.dip_positions.shape[0]


@dataclass
class RegularGrid:
    """engine_grid.RegularGrid(orthogonal_extent, regular_grid_shape) (_engine_factory.py:72-75,93-96).

    Cell centres, ``meshgrid(indexing='ij')`` order: x slowest, z fastest
    (gempy/core/data/grid_modules/regular_grid.py:58-71,191-201).  ``values`` is generated lazily;
    the CUDA path never needs it for a regular grid (coordinates come from the linear index)."""
    orthogonal_extent: np.ndarray
    regular_grid_shape: np.ndarray
    _values: object = field(default=None, repr=False)   # explicit centres (octree levels > 0); may be Deferred
    _dxdydz: Optional[np.ndarray] = field(default=None, repr=False)
    _n_points: Optional[int] = field(default=None, repr=False)

    def __post_init__(self):
        self.orthogonal_extent = np.asarray(self.orthogonal_extent, dtype=np.float64).reshape(6)
        self.regular_grid_shape = np.asarray(self.regular_grid_shape, dtype=np.int64).reshape(3)

    @classmethod
    def from_octree_level(cls, xyz_coords_octree: np.ndarray, previous_regular_grid: "RegularGrid",
                          active_cells=None, left_right=None) -> "RegularGrid":
        g = cls(previous_regular_grid.orthogonal_extent, previous_regular_grid.regular_grid_shape * 2)
        g._values = xyz_coords_octree if isinstance(xyz_coords_octree, Deferred) else \
            np.ascontiguousarray(xyz_coords_octree, dtype=np.float64)
        g._dxdydz = previous_regular_grid.dxdydz / 2.0
        return g

    @property
    def is_implicit(self) -> bool:
        return self._values is None

    @property
    def dxdydz(self) -> np.ndarray:
        if self._dxdydz is not None:
            return self._dxdydz
        e, s = self.orthogonal_extent, self.regular_grid_shape
        return np.array([(e[1] - e[0]) / s[0], (e[3] - e[2]) / s[1], (e[5] - e[4]) / s[2]])

    @property
    def resolution(self) -> np.ndarray:
        return self.regular_grid_shape

    def axis_coords(self):
        e, s = self.orthogonal_extent, self.regular_grid_shape
        d = self.dxdydz if self._dxdydz is None else np.array(
            [(e[1] - e[0]) / s[0], (e[3] - e[2]) / s[1], (e[5] - e[4]) / s[2]])
        return [np.linspace(e[2 * a] + d[a] / 2, e[2 * a + 1] - d[a] / 2, int(s[a]), dtype=np.float64)
                for a in range(3)]

    @property
    def values(self) -> np.ndarray:
        if self._values is None:
            g = np.meshgrid(*self.axis_coords(), indexing="ij")
            self._values = np.vstack([c.ravel() for c in g]).T.astype(np.float64)
        self._values = _res(self._values)
        return self._values

    @property
    def n_points(self) -> int:
        if self._n_points is not None:
            return self._n_points
        if self._values is not None:
            return _res(self._values).shape[0]
        return int(np.prod(self.regular_grid_shape))


@dataclass
class GenericGrid:
    """engine_grid.GenericGrid(values) (_engine_factory.py:76-81)."""
    values: np.ndarray

    def __post_init__(self):
        self.values = np.ascontiguousarray(self.values, dtype=np.float64).reshape(-1, 3)

    @property
    def n_points(self) -> int:
        return self.values.shape[0]


@dataclass
class CenteredGrid:
    """engine_grid.CenteredGrid(centers, radius, resolution) (_engine_factory.py:82-87): the voxelised kernel of the
    forward-gravity computation, repeated around every device centre."""
    centers: np.ndarray
    radius: np.ndarray
    resolution: np.ndarray

    def __post_init__(self):
        from .geophysics import centered_grid_kernel
        self.centers = np.ascontiguousarray(self.centers, dtype=np.float64).reshape(-1, 3)
        self.radius = np.broadcast_to(np.asarray(self.radius, dtype=np.float64).ravel(), (3,)).copy()
        self.resolution = np.asarray(self.resolution, dtype=np.int64).reshape(3)
        self.kernel_grid_centers, self.kernel_dxyz_left, self.kernel_dxyz_right = centered_grid_kernel(
            self.resolution, self.radius)

    @property
    def values(self) -> np.ndarray:
        return (self.centers[:, None, :] + self.kernel_grid_centers[None, :, :]).reshape(-1, 3)

    @property
    def n_points(self) -> int:
        return self.centers.shape[0] * self.kernel_grid_centers.shape[0]


@dataclass
class EngineGrid:
    """engine_grid.EngineGrid (_engine_factory.py:97-104).  Evaluation order of the point sets is
    octree, dense, custom, topography, sections, (geophysics)."""
    octree_grid: Optional[RegularGrid] = None
    dense_grid: Optional[RegularGrid] = None
    topography: Optional[GenericGrid] = None
    sections: Optional[GenericGrid] = None
    custom_grid: Optional[GenericGrid] = None
    geophysics_grid: Optional[CenteredGrid] = None

    _ORDER = ("octree_grid", "dense_grid", "custom_grid", "topography", "sections", "geophysics_grid")

    def parts(self):
        """[(name, grid)] of the active point sets in evaluation order."""
        return [(n, getattr(self, n)) for n in self._ORDER if getattr(self, n) is not None]

    def _slice_of(self, name: str) -> slice:
        start = 0
        for n, g in self.parts():
            if n == name:
                return slice(start, start + g.n_points)
            start += g.n_points
        return slice(0, 0)

    octree_grid_slice = property(lambda self: self._slice_of("octree_grid"))
    dense_grid_slice = property(lambda self: self._slice_of("dense_grid"))
    custom_grid_slice = property(lambda self: self._slice_of("custom_grid"))
    topography_slice = property(lambda self: self._slice_of("topography"))
    sections_slice = property(lambda self: self._slice_of("sections"))
    geophysics_grid_slice = property(lambda self: self._slice_of("geophysics_grid"))

    @property
    def len_all_grids(self) -> int:
        return sum(g.n_points for _, g in self.parts())

    @property
    def values(self) -> np.ndarray:
        return np.vstack([g.values for _, g in self.parts()]) if self.parts() else np.zeros((0, 3))

    @classmethod
    def from_xyz_coords(cls, xyz_coords: np.ndarray) -> "EngineGrid":
        return cls(custom_grid=GenericGrid(xyz_coords))


@dataclass
class InterpolationInput:
    """InterpolationInput(surface_points, orientations, grid, unit_values, weights)
    (_engine_factory.py:50-56)."""
    surface_points: SurfacePoints
    orientations: Orientations
    grid: EngineGrid
    unit_values: Optional[np.ndarray] = None
    weights: Optional[List[np.ndarray]] = None

    def __post_init__(self):
        if self.weights is None:
            self.weights = []


# --------------------------------------------------------------------------- descriptor
@dataclass
class FaultsData:
    """Placeholder for finite-fault data (gempy/core/data/structural_group.py:32).  Finite faults
    are a prototype in the reference (gempy/API/faults_API.py:84-91 raises NotImplementedError);
    this backend rejects them."""
    fault_values_everywhere: Optional[np.ndarray] = None
    fault_values_on_sp: Optional[np.ndarray] = None
    thickness: Optional[float] = None

    @property
    def finite_faults_defined(self) -> bool:
        return self.thickness is not None


@dataclass
class TensorsStructure:
    number_of_points_per_surface: np.ndarray

    def __post_init__(self):
        self.number_of_points_per_surface = np.asarray(self.number_of_points_per_surface, dtype=np.int64)

    @property
    def n_surfaces(self) -> int:
        return int(self.number_of_points_per_surface.shape[0])

    @property
    def total_number_sp(self) -> int:
        return int(self.number_of_points_per_surface.sum())

    @property
    def reference_sp_position(self) -> np.ndarray:
        """Index of the reference point of every surface = its first point."""
        n = self.number_of_points_per_surface
        return np.concatenate([[0], np.cumsum(n)[:-1]]).astype(np.int64)


@dataclass
class StacksStructure:
    number_of_points_per_stack: np.ndarray
    number_of_orientations_per_stack: np.ndarray
    number_of_surfaces_per_stack: np.ndarray
    masking_descriptor: Sequence[StackRelationType]
    faults_relations: Optional[np.ndarray] = None
    faults_input_data: Optional[List[Optional[FaultsData]]] = None

    def __post_init__(self):
        self.number_of_points_per_stack = np.asarray(self.number_of_points_per_stack, dtype=np.int64)
        self.number_of_orientations_per_stack = np.asarray(self.number_of_orientations_per_stack, dtype=np.int64)
        self.number_of_surfaces_per_stack = np.asarray(self.number_of_surfaces_per_stack, dtype=np.int64)
        if self.faults_relations is not None:
            self.faults_relations = np.asarray(self.faults_relations, dtype=bool)

    @property
    def n_stacks(self) -> int:
        return int(self.number_of_points_per_stack.shape[0])


@dataclass
class InputDataDescriptor:
    """InputDataDescriptor.from_structural_frame(structural_frame, making_descriptor,
    faults_relations, faults_input_data) (gempy/core/data/structural_frame.py:311-316)."""
    tensors_structure: TensorsStructure
    stack_structure: StacksStructure

    @classmethod
    def from_structural_frame(cls, structural_frame, making_descriptor, faults_relations,
                              faults_input_data=None) -> "InputDataDescriptor":
        # reads what gempy/core/data/structural_frame.py:333-350 exposes.  structural_elements ends with the basement
        # element (no surface points: number_of_points_per_element = [..., 0]) while number_of_elements_per_group does not
        # count it: only the elements that belong to a group describe surfaces.
        n_surfaces = int(np.sum(structural_frame.number_of_elements_per_group))
        ts = TensorsStructure(np.asarray(structural_frame.number_of_points_per_element)[:n_surfaces])
        ss = StacksStructure(
            number_of_points_per_stack=structural_frame.number_of_points_per_group,
            number_of_orientations_per_stack=structural_frame.number_of_orientations_per_group,
            number_of_surfaces_per_stack=structural_frame.number_of_elements_per_group,
            masking_descriptor=list(making_descriptor),
            faults_relations=faults_relations,
            faults_input_data=faults_input_data,
        )
        return cls(ts, ss)


# --------------------------------------------------------------------------- options
@dataclass
class KernelOptions:
    """Defaults pinned by the serialization golden
    test/test_modules/test_serialize_model.test_generate_horizontal_stratigraphic_model.verify/
    'Horizontal Stratigraphic Model serialization.approved.txt' (kernel_options block)."""
    range: float = 1.7
    c_o: float = 10.0
    uni_degree: int = 1
    i_res: float = 4.0
    gi_res: float = 2.0
    number_dimensions: int = 3
    kernel_function: AvailableKernelFunctions = AvailableKernelFunctions.cubic
    kernel_solver: int = 1                      # 1 = direct dense solve
    compute_condition_number: bool = False
    optimizing_condition_number: bool = False
    condition_number: Optional[float] = None


@dataclass
class EvaluationOptions:
    """Same golden, evaluation_options block."""
    _number_octree_levels: int = 1
    _number_octree_levels_surface: int = 4
    octree_curvature_threshold: float = -1.0
    octree_error_threshold: float = 1.0
    octree_min_level: int = 2
    mesh_extraction: bool = True
    mesh_extraction_masking_options: MeshExtractionMaskingOptions = MeshExtractionMaskingOptions.INTERSECT
    mesh_extraction_fancy: bool = True
    evaluation_chunk_size: int = 500_000
    compute_scalar_gradient: bool = False
    verbose: bool = False

    @property
    def number_octree_levels(self) -> int:
        return self._number_octree_levels

    @number_octree_levels.setter
    def number_octree_levels(self, v: int):
        self._number_octree_levels = int(v)

    @property
    def number_octree_levels_surface(self) -> int:
        return min(self._number_octree_levels_surface, self._number_octree_levels)

    @number_octree_levels_surface.setter
    def number_octree_levels_surface(self, v: int):
        self._number_octree_levels_surface = int(v)


@dataclass
class InterpolationOptions:
    kernel_options: KernelOptions = field(default_factory=KernelOptions)
    evaluation_options: EvaluationOptions = field(default_factory=EvaluationOptions)
    sigmoid_slope: float = 5_000_000.0
    debug: bool = False
    cache_mode: int = 3
    cache_model_name: str = ""
    block_solutions_type: BlockSolutionType = BlockSolutionType.OCTREE

    # constructors the reference calls (gempy/API/initialization_API.py:81-83,
    # gempy/modules/json_io/json_operations.py:158)
    @classmethod
    def from_args(cls, range: float = 1.7, c_o: float = 10.0, uni_degree: int = 1, i_res: float = 4.0,
                  gi_res: float = 2.0, number_dimensions: int = 3, number_octree_levels: int = 1,
                  kernel_function: AvailableKernelFunctions = AvailableKernelFunctions.cubic,
                  mesh_extraction: bool = True, compute_scalar_gradient: bool = False,
                  sigmoid_slope: float = 5_000_000.0) -> "InterpolationOptions":
        ko = KernelOptions(range=range, c_o=c_o, uni_degree=uni_degree, i_res=i_res, gi_res=gi_res,
                           number_dimensions=number_dimensions, kernel_function=kernel_function)
        eo = EvaluationOptions(_number_octree_levels=number_octree_levels, mesh_extraction=mesh_extraction,
                               compute_scalar_gradient=compute_scalar_gradient)
        return cls(kernel_options=ko, evaluation_options=eo, sigmoid_slope=sigmoid_slope)

    @classmethod
    def init_octree_options(cls, range: float = 1.7, c_o: float = 10.0, refinement: int = 1) -> "InterpolationOptions":
        return cls.from_args(range=range, c_o=c_o, number_octree_levels=refinement, mesh_extraction=True)

    @classmethod
    def init_dense_grid_options(cls) -> "InterpolationOptions":
        o = cls.from_args(number_octree_levels=1, mesh_extraction=False)
        o.block_solutions_type = BlockSolutionType.DENSE_GRID
        return o

    # shortcuts the reference's tests/examples touch
    # (test_custom_grid.py:30, Alesmodel.py:107-145, Moureze.py:151-154)
    @property
    def number_octree_levels(self) -> int:
        return self.evaluation_options.number_octree_levels

    @number_octree_levels.setter
    def number_octree_levels(self, v: int):
        self.evaluation_options.number_octree_levels = v

    @property
    def number_octree_levels_surface(self) -> int:
        return self.evaluation_options.number_octree_levels_surface

    @number_octree_levels_surface.setter
    def number_octree_levels_surface(self, v: int):
        self.evaluation_options.number_octree_levels_surface = v

    @property
    def mesh_extraction(self) -> bool:
        return self.evaluation_options.mesh_extraction

    @mesh_extraction.setter
    def mesh_extraction(self, v: bool):
        self.evaluation_options.mesh_extraction = bool(v)

    @property
    def compute_scalar_gradient(self) -> bool:
        return self.evaluation_options.compute_scalar_gradient

    @compute_scalar_gradient.setter
    def compute_scalar_gradient(self, v: bool):
        self.evaluation_options.compute_scalar_gradient = bool(v)


# --------------------------------------------------------------------------- transform
class GlobalAnisotropy(enum.Enum):
    """Anisotropy policy of the input transform (gempy/core/data/geo_model.py:274-279,
    gempy/API/examples_generator.py:60,551)."""
    CUBE = enum.auto()       # rescale every axis to the unit cube
    NONE = enum.auto()       # one isotropic scale (the reference examples used here all end up with this)
    MANUAL = enum.auto()


@dataclass
class Transform:
    """Input transform ``x' = (x + position) * scale`` (gempy/core/data/geo_model.py:135-141,158-168,244-247; rotations are
    always 0 in the reference's models and unsupported here).  The golden JSON pins HORIZONTAL_STRAT to position [-500]*3,
    scale 6.25e-4; ``from_input_points`` reproduces the transform stored in the reference's Greenstone.gempy bit for bit."""
    position: np.ndarray
    rotation: np.ndarray
    scale: np.ndarray
    _is_default_transform: bool = False
    _cached_pivot: Optional[np.ndarray] = None

    def __post_init__(self):
        self.position = np.asarray(self.position, dtype=np.float64).reshape(3)
        self.rotation = np.asarray(self.rotation, dtype=np.float64).reshape(3)
        self.scale = np.broadcast_to(np.asarray(self.scale, dtype=np.float64).ravel(), (3,)).copy()

    @classmethod
    def init_neutral(cls) -> "Transform":
        return cls(position=np.zeros(3), rotation=np.zeros(3), scale=np.ones(3), _is_default_transform=True)

    @classmethod
    def from_input_points(cls, surface_points, orientations) -> "Transform":
        """Accepts coordinate arrays or the reference's SurfacePointsTable / OrientationsTable (``.xyz``)."""
        sp = np.asarray(getattr(surface_points, "xyz", surface_points), dtype=np.float64).reshape(-1, 3)
        op = np.asarray(getattr(orientations, "xyz", orientations), dtype=np.float64).reshape(-1, 3)
        pts = np.concatenate([sp, op], axis=0)
        mx, mn = pts.max(axis=0), pts.min(axis=0)
        scaling = 2.0 * np.max(mx - mn)
        center = (mx + mn) / 2.0
        f = 1.0 / scaling
        return cls(position=-center, rotation=np.zeros(3), scale=np.array([f, f, f]))

    @classmethod
    def from_matrix(cls, matrix: np.ndarray) -> "Transform":
        m = np.asarray(matrix, dtype=np.float64)
        if not np.allclose(m[:3, :3], np.diag(np.diag(m[:3, :3]))):
            raise NotImplementedError("rotated grids are outside the B200 backend's scope")
        return cls(position=m[:3, 3].copy(), rotation=np.zeros(3), scale=np.diag(m[:3, :3]).copy())

    def _no_rotation(self):
        if np.any(self.rotation != 0):
            raise NotImplementedError("rotated transforms are outside the B200 backend's scope")

    def apply(self, points: np.ndarray) -> np.ndarray:
        self._no_rotation()
        return (np.asarray(points, dtype=np.float64) + self.position) * self.scale

    def apply_inverse(self, points: np.ndarray) -> np.ndarray:
        self._no_rotation()
        return np.asarray(points, dtype=np.float64) / self.scale - self.position

    # pivot variants: a rotation-free transform does not depend on the pivot
    def apply_with_pivot(self, points: np.ndarray, pivot=None) -> np.ndarray:
        return self.apply(points)

    def apply_inverse_with_pivot(self, points: np.ndarray, pivot=None) -> np.ndarray:
        return self.apply_inverse(points)

    def apply_with_cached_pivot(self, points: np.ndarray) -> np.ndarray:
        return self.apply(points)

    def apply_inverse_with_cached_pivot(self, points: np.ndarray) -> np.ndarray:
        return self.apply_inverse(points)

    def apply_anisotropy(self, anisotropy_type: "GlobalAnisotropy" = GlobalAnisotropy.NONE, anisotropy_limit=None) -> None:
        if anisotropy_type in (GlobalAnisotropy.NONE, None):
            self.scale = np.full(3, float(np.min(self.scale)) if np.ptp(self.scale) else float(self.scale[0]))
            return
        raise NotImplementedError("only GlobalAnisotropy.NONE is supported by the B200 backend's transform")

    def __add__(self, other: "Transform") -> "Transform":
        return Transform(position=self.position + other.position, rotation=self.rotation + other.rotation,
                         scale=self.scale * other.scale)

    def transform_gradient(self, gradients: np.ndarray) -> np.ndarray:
        g = np.asarray(gradients, dtype=np.float64)
        t = g / self.scale                      # inverse-transpose of diag(scale)
        n0 = np.linalg.norm(g, axis=1)
        n1 = np.linalg.norm(t, axis=1)
        n1[n1 == 0] = 1.0
        return t * (n0 / n1)[:, None]

    def scale_points(self, points: np.ndarray) -> np.ndarray:
        return np.asarray(points, dtype=np.float64) * self.scale

    def get_matrix_4x4(self) -> np.ndarray:
        m = np.eye(4)
        m[:3, :3] = np.diag(self.scale)
        m[:3, 3] = self.position * self.scale
        return m


# --------------------------------------------------------------------------- outputs
@dataclass
class ExportedFields:
    """Scalar field (and optional gradient) on [all grid points ++ all surface points of the stack's
    model]; the ``scalar_field`` view drops the surface-point tail
    (test/test_model_types/test_example_models_I.py:20-21)."""
    _scalar_field: object
    _gx_field: object = None
    _gy_field: object = None
    _gz_field: object = None
    _grid_size: int = 0
    _scalar_field_at_surface_points: object = None

    @property
    def scalar_field_everywhere(self) -> np.ndarray:
        self._scalar_field = _res(self._scalar_field)
        return self._scalar_field

    @property
    def scalar_field(self) -> np.ndarray:
        return self.scalar_field_everywhere[:self._grid_size]

    def _g(self, name):
        v = _res(getattr(self, name))
        setattr(self, name, v)
        return v

    @property
    def gx_field(self):
        v = self._g("_gx_field")
        return None if v is None else v[:self._grid_size]

    @property
    def gy_field(self):
        v = self._g("_gy_field")
        return None if v is None else v[:self._grid_size]

    @property
    def gz_field(self):
        v = self._g("_gz_field")
        return None if v is None else v[:self._grid_size]

    @property
    def scalar_field_at_surface_points(self) -> np.ndarray:
        self._scalar_field_at_surface_points = _res(self._scalar_field_at_surface_points)
        return self._scalar_field_at_surface_points


@dataclass
class ScalarFieldOutput:
    _weights: object
    grid: EngineGrid
    exported_fields: ExportedFields
    _values_block: object               # (1, n_xyz) activator output on grid ++ surface points
    stack_relation: StackRelationType
    _mask_components: object = None

    @property
    def weights(self) -> np.ndarray:
        self._weights = _res(self._weights)
        return self._weights

    @property
    def values_block(self) -> np.ndarray:
        self._values_block = _res(self._values_block)
        return self._values_block

    @property
    def mask_components(self) -> np.ndarray:
        self._mask_components = _res(self._mask_components)
        return self._mask_components

    @property
    def grid_size(self) -> int:
        return self.exported_fields._grid_size


@dataclass
class CombinedScalarFieldsOutput:
    _squeezed_mask_array: object        # (n_xyz,) bool: where this stack owns the final block
    _final_block: object                # (n_xyz,) combined lith block (same for every stack)
    _faults_block: object               # (n_xyz,) sum of fault blocks
    final_exported_fields: Optional[ExportedFields] = None

    @property
    def squeezed_mask_array(self) -> np.ndarray:
        self._squeezed_mask_array = _res(self._squeezed_mask_array)
        return self._squeezed_mask_array

    @property
    def final_block(self) -> np.ndarray:
        self._final_block = _res(self._final_block)
        return self._final_block

    @property
    def faults_block(self) -> np.ndarray:
        self._faults_block = _res(self._faults_block)
        return self._faults_block


@dataclass
class InterpOutput:
    scalar_fields: ScalarFieldOutput
    combined_scalar_field: Optional[CombinedScalarFieldsOutput] = None

    @property
    def weights(self):
        return self.scalar_fields.weights

    @property
    def grid(self) -> EngineGrid:
        return self.scalar_fields.grid

    @property
    def exported_fields(self) -> ExportedFields:
        return self.scalar_fields.exported_fields

    @property
    def exported_fields_dense_grid(self) -> ExportedFields:
        sl = self.grid.dense_grid_slice
        ef = self.scalar_fields.exported_fields
        pick = lambda a: None if a is None else a[sl]
        return ExportedFields(pick(ef.scalar_field_everywhere), pick(ef._g("_gx_field")), pick(ef._g("_gy_field")),
                              pick(ef._g("_gz_field")), sl.stop - sl.start, ef.scalar_field_at_surface_points)

    @property
    def values_block(self) -> np.ndarray:
        return self.scalar_fields.values_block[:, :self.scalar_fields.grid_size]

    @property
    def block(self) -> np.ndarray:
        return self.combined_scalar_field.final_block[:self.scalar_fields.grid_size]

    @property
    def ids_block(self) -> np.ndarray:
        return np.rint(self.block)

    @property
    def faults_block(self) -> np.ndarray:
        return self.combined_scalar_field.faults_block[:self.scalar_fields.grid_size]

    @property
    def litho_faults_ids(self) -> np.ndarray:
        lith = np.rint(self.block)
        faults = np.rint(self.faults_block)
        mult = max(int(len(np.unique(lith))), 1)
        return lith + faults * mult


@dataclass
class OctreeLevel:
    grid_centers: EngineGrid
    outputs_centers: List[InterpOutput]
    _grid_corners: object = None
    outputs_corners: Optional[List[InterpOutput]] = None
    _marked_voxels: object = None                   # refine mask over this level's voxels (lazy)
    _device_fields: object = None                   # FieldsOnDevice of this level (device consumers: marching cubes)

    @property
    def marked_voxels(self) -> Optional[np.ndarray]:
        self._marked_voxels = _res(self._marked_voxels)
        return self._marked_voxels

    @property
    def grid_corners(self) -> Optional[EngineGrid]:
        self._grid_corners = _res(self._grid_corners)
        return self._grid_corners

    @property
    def outputs(self) -> List[InterpOutput]:
        return self.outputs_centers

    @property
    def last_output_center(self) -> InterpOutput:
        return self.outputs_centers[-1]

    @property
    def number_of_outputs(self) -> int:
        return len(self.outputs_centers)

    @property
    def dxdydz(self):
        return self.grid_centers.octree_grid.dxdydz


@dataclass
class DualContouringData:
    xyz_on_edge: np.ndarray
    valid_edges: np.ndarray
    gradients: Optional[np.ndarray] = None


class DualContouringMesh:
    """``dc_meshes[e]`` of the engine's Solutions: ``vertices`` (V, 3) float64, ``edges`` (T, 3) int64 triangles, ``dc_data``
    (gempy/core/data/geo_model.py:110-121 reads the first two).  The B200 path hands in ``Deferred`` values: the mesh stays
    on the device until an attribute is read."""

    def __init__(self, vertices, edges, dc_data=None):
        self._vertices = vertices
        self._edges = edges
        self._dc_data = dc_data

    @property
    def vertices(self) -> np.ndarray:
        self._vertices = _res(self._vertices)
        return self._vertices

    @vertices.setter
    def vertices(self, v):
        self._vertices = v

    @property
    def edges(self) -> np.ndarray:
        self._edges = _res(self._edges)
        return self._edges

    @edges.setter
    def edges(self, v):
        self._edges = v

    @property
    def dc_data(self) -> Optional[DualContouringData]:
        self._dc_data = _res(self._dc_data)
        return self._dc_data

    @property
    def vertices_tensor(self):
        return self.vertices


class RawArraysSolution:
    """Dense arrays GemPy users read after ``compute_model``:
    lith_block, fault_block, litho_faults_block, scalar_field_matrix, block_matrix, mask_matrix,
    mask_matrix_squeezed, custom, vertices, edges.  Each array is materialised on first access."""
    BlockSolutionType = BlockSolutionType
    _DEFAULTS = {
        "lith_block": lambda: np.empty(0), "fault_block": lambda: np.empty(0), "litho_faults_block": lambda: np.empty(0),
        "scalar_field_matrix": lambda: np.empty((0, 0)), "block_matrix": lambda: np.empty((0, 0)),
        "mask_matrix": lambda: np.empty((0, 0)), "mask_matrix_squeezed": lambda: np.empty((0, 0)),
        "custom": lambda: None, "topography": lambda: None, "sections": lambda: None, "dense_ids": lambda: None,
        "vertices": list, "edges": list,
    }

    def __init__(self):
        self._lazy = {}

    def set_lazy(self, name: str, fn) -> None:
        self._lazy[name] = fn

    def __getattr__(self, name):            # reached only when the attribute is not materialised yet
        lazy = self.__dict__.get("_lazy", {})
        if name in lazy:
            v = lazy.pop(name)()
        elif name in RawArraysSolution._DEFAULTS:
            v = RawArraysSolution._DEFAULTS[name]()
        else:
            raise AttributeError(name)
        setattr(self, name, v)
        return v


class Solutions:
    """What ``compute_model`` returns (gempy/API/compute_API.py:68-73) and
    ``GeoModel.solutions`` consumes (gempy/core/data/geo_model.py:100-127)."""

    def __init__(self, octrees_output: List[OctreeLevel], dc_meshes: Optional[List[DualContouringMesh]] = None,
                 fw_gravity=None, block_solution_type: BlockSolutionType = BlockSolutionType.OCTREE):
        self.octrees_output = octrees_output
        self.dc_meshes = dc_meshes
        self.gravity = fw_gravity
        self.block_solution_type = block_solution_type
        self.scalar_field_at_surface_points: List[float] = []
        self._ordered_elements: List[np.ndarray] = []
        for out in octrees_output[0].outputs_centers:
            sfasp = out.exported_fields.scalar_field_at_surface_points
            self.scalar_field_at_surface_points.extend(np.asarray(sfasp).tolist())
            self._ordered_elements.append(np.argsort(sfasp)[::-1])
        self.raw_arrays: Optional[RawArraysSolution] = None

    @property
    def root_output(self) -> OctreeLevel:
        return self.octrees_output[0]

    def meshes_to_unstruct(self):
        raise NotImplementedError("subsurface export is outside the B200 backend's scope (SURVEY.md §2 row 13)")
