"""Dense-grid mesh extraction by marching cubes on the device.

Mirror of ``gempy/modules/mesh_extranction/marching_cubes.py``: ``set_meshes_with_marching_cubes(model)`` (lines 13-55)
walks the structural groups, takes each group's scalar field on the dense grid, the squeezed mask of the group (none for
faults) and each element's isovalue, and stores ``vertices`` / ``edges`` on the element (lines 82-101).  The reference
hands the arrays to ``skimage.measure.marching_cubes``; here they never leave the GPU: ``gpb_mc_count`` /
``gpb_mc_emit`` work on the level-0 device fields that ``compute_model`` keeps behind the lazy ``Solutions``.

Conventions kept from the reference: a cube is meshed when the mask is set at its far corner (pinned by the vertex counts
of test/test_modules/test_marching_cubes.py:44-47), vertices are ``index * (dx, dy, dz) + extent minima`` in real
coordinates (the half-cell offset of the cell centres is not added, marching_cubes.py:92-95)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .data import BlockSolutionType, DualContouringMesh, Solutions, StackRelationType


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def marching_cubes_device(Z: torch.Tensor, shape: Sequence[int], level: float, mask: Optional[torch.Tensor] = None,
                          spacing: Sequence[float] = (1.0, 1.0, 1.0), origin: Sequence[float] = (0.0, 0.0, 0.0),
                          triangles: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Z: contiguous float64 CUDA tensor of nx*ny*nz values (x slowest, z fastest); mask: uint8/bool tensor of the same
    length or None.  Returns device tensors: vertices [V, 3] float64, triangles [T, 3] int32."""
    lib = _lib.lib()
    if not Z.is_cuda:
        raise _lib.GpbError("marching cubes runs on the device fields (no CPU fallback)")
    nx, ny, nz = (int(v) for v in shape)
    m = nx * ny * nz
    if Z.numel() != m or Z.dtype != torch.float64 or not Z.is_contiguous():
        raise ValueError("Z must be a contiguous float64 tensor of nx*ny*nz values")
    if mask is not None:
        mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
        if mask.numel() != m or mask.dtype != torch.uint8 or not mask.is_contiguous():
            raise ValueError("mask must be a contiguous uint8/bool tensor of nx*ny*nz values")
    dev = Z.device
    stream = torch.cuda.current_stream(dev).cuda_stream
    flags = torch.empty(m, dtype=torch.uint8, device=dev)
    offsets = torch.empty(int(lib.gpb_mc_scratch_elems(m)), dtype=torch.int64, device=dev)
    nv, nt = C.c_longlong(), C.c_longlong()
    with torch.cuda.device(dev):
        _lib.check(lib.gpb_mc_count(_ptr(Z), _ptr(mask), nx, ny, nz, float(level), _ptr(flags), _ptr(offsets),
                                    C.byref(nv), C.byref(nt), stream))
        verts = torch.empty((nv.value, 3), dtype=torch.float64, device=dev)
        tris = torch.empty((nt.value, 3), dtype=torch.int32, device=dev) if triangles else None
        if nv.value:
            vbase = torch.empty(m, dtype=torch.int32, device=dev)
            _lib.check(lib.gpb_mc_emit(_ptr(Z), _ptr(flags), _ptr(offsets), nx, ny, nz, float(level),
                                       float(origin[0]), float(origin[1]), float(origin[2]),
                                       float(spacing[0]), float(spacing[1]), float(spacing[2]),
                                       _ptr(vbase), _ptr(verts), _ptr(tris) if nt.value and triangles else None, stream))
    return verts, tris


def extract_meshes(solutions: Solutions, extent: Sequence[float], resolution: Sequence[int]) -> List[DualContouringMesh]:
    """One mesh per (stack, surface) in stack order, vertices in real coordinates.  ``extent``/``resolution`` are the
    real-coordinate dense grid (``model.grid.regular_grid`` in the reference, marching_cubes.py:30)."""
    if solutions is None or solutions.block_solution_type != BlockSolutionType.DENSE_GRID:
        raise ValueError("Model solutions must contain dense grid data for mesh extraction.")     # marching_cubes.py:27-28
    if not solutions.octrees_output or not solutions.octrees_output[0].outputs:
        raise ValueError("No interpolation outputs available for mesh extraction.")               # marching_cubes.py:33-34
    lvl0 = solutions.octrees_output[0]
    f = getattr(lvl0, "_device_fields", None)
    if f is None:
        raise _lib.GpbError("these solutions carry no device fields (not produced by the B200 backend)")
    shape = np.asarray(resolution, dtype=int)
    ext = np.asarray(extent, dtype=float)
    spacing = (ext[1::2] - ext[0::2]) / shape
    sl = f.seg_slice("dense_grid")
    if sl.stop - sl.start != int(np.prod(shape)):
        raise ValueError("resolution does not match the dense grid the model was computed on")
    meshes = []
    for i, out in enumerate(lvl0.outputs):
        Z = f.Z[i, sl].contiguous()
        is_fault = out.scalar_fields.stack_relation == StackRelationType.FAULT
        mask = None if is_fault else f.squeezed[i, sl].contiguous()
        for iso in f.isovalues[i].cpu().numpy().tolist():
            v, t = marching_cubes_device(Z, shape, iso, mask, spacing, ext[0::2])
            meshes.append(DualContouringMesh(v.cpu().numpy(), t.cpu().numpy().astype(np.int64)))
    return meshes


def set_meshes_with_marching_cubes(model) -> None:
    """Same contract as the reference function (marching_cubes.py:13-55): reads ``model.solutions``,
    ``model.grid.regular_grid`` and ``model.structural_frame.structural_groups``; writes ``vertices``/``edges`` on
    every structural element.  Like the reference (marching_cubes.py:36-47) every element gets the mesh of ITS OWN
    isovalue, ``element.scalar_field_at_interface``: the ``GeoModel.solutions`` setter has already reordered each group's
    elements by decreasing scalar value (geo_model.py:121-127), so pairing meshes and elements by position would hand
    elements their neighbours' surfaces whenever that reorder was not the identity."""
    rg = model.grid.regular_grid
    sol = model.solutions
    meshes = extract_meshes(sol, rg.extent, rg.resolution)
    f = sol.octrees_output[0]._device_fields
    n_out = len(sol.octrees_output[0].outputs)
    k = 0
    for e, group in enumerate(model.structural_frame.structural_groups):
        if e >= n_out:
            continue
        iso = f.isovalues[e].cpu().numpy()
        group_meshes = meshes[k:k + iso.shape[0]]
        k += iso.shape[0]
        for pos, element in enumerate(group.elements):
            level = getattr(element, "scalar_field_at_interface", None)
            j = pos if level is None else int(np.argmin(np.abs(iso - float(level))))
            if j < len(group_meshes):
                element.vertices = group_meshes[j].vertices
                element.edges = group_meshes[j].edges
