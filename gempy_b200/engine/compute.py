"""Host side of the B200 backend: ``compute_model`` with the engine's signature.

Mirrors ``gempy_engine.compute_model(interpolation_input, options, data_descriptor, geophysics_input)``
(the call GemPy makes at /root/reference/gempy/API/compute_API.py:68-73) and returns a ``Solutions`` object
with the attributes GemPy reads (gempy/core/data/geo_model.py:100-127).  The pipeline per octree level is

    per stack:  ref/rest split -> [assemble covariance -> LU solve | cached weights] -> pack weights
                -> fused field(+gradient) evaluation on {octree centres, dense, custom, topography, sections,
                   corners} ++ surface points -> activator
    all stacks: masks + combination -> lith / fault blocks
    next level: corner-id refinement test -> child emission
    finally:    dual contouring on the surface level

Every arithmetic step is a CUDA kernel of libgempy_b200.so called through the C ABI (gempy_b200/_lib.py).
torch is used for device-memory ownership, streams and (multi-GPU) torch.distributed only.
There is no CPU fallback: without the library or a CUDA device this module raises.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .comm import Comm
from .data import (BlockSolutionType, CombinedScalarFieldsOutput, Deferred, DualContouringData, DualContouringMesh, EngineGrid,
                   ExportedFields, GenericGrid, InputDataDescriptor, InterpolationInput, InterpolationOptions,
                   InterpOutput, OctreeLevel, RawArraysSolution, RegularGrid, ScalarFieldOutput, Solutions,
                   StackRelationType)

GRID_SHIFT = 1e-6          # regular-grid / octree points sit at centre + 1e-6 (pinned by the approved vectors)
F64 = torch.float64


def _rel_code(rel) -> int:
    if rel is False or rel is None:
        return StackRelationType.BASEMENT.value
    return int(getattr(rel, "value", rel))


def _kernel_code(k) -> int:
    return _lib.GPB_KERNEL[getattr(k, "name", k)]


def _n_drift(degree: int) -> int:
    return {0: 0, 1: 3, 2: 9}[int(degree)]


def _ptr(t: Optional[torch.Tensor], byte_offset: int = 0) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr() + byte_offset


# ------------------------------------------------------------------------------------------------ segments
@dataclass
class Segment:
    """A set of evaluation points: either an implicit regular grid range or an explicit [3, m] device table."""
    name: str
    m: int
    grid: Optional[_lib.GpbRegularGrid] = None
    i0: int = 0
    xyz: Optional[torch.Tensor] = None          # [3, m] contiguous


def regular_descriptor(g: RegularGrid) -> _lib.GpbRegularGrid:
    e, s = g.orthogonal_extent, g.regular_grid_shape
    d = np.array([(e[1] - e[0]) / s[0], (e[3] - e[2]) / s[1], (e[5] - e[4]) / s[2]])
    return _lib.GpbRegularGrid(e[0] + d[0] / 2 + GRID_SHIFT, e[2] + d[1] / 2 + GRID_SHIFT, e[4] + d[2] / 2 + GRID_SHIFT,
                               d[0], d[1], d[2], int(s[0]), int(s[1]), int(s[2]))


def regular_centers_device(g: RegularGrid, device) -> Tuple[torch.Tensor, np.ndarray]:
    """Explicit [3, n] centres of a (small) regular grid, x slowest / z fastest, shift included."""
    ax = g.axis_coords()
    gx, gy, gz = np.meshgrid(*ax, indexing="ij")
    xyz = np.stack([gx.ravel(), gy.ravel(), gz.ravel()]) + GRID_SHIFT
    return torch.as_tensor(xyz, dtype=F64, device=device).contiguous(), g.dxdydz.copy()


# ------------------------------------------------------------------------------------------------ stack tables
class StackTables:
    """Device tables of one stack (the ref/rest split of the engine's preprocess stage)."""

    def __init__(self, ii: InterpolationInput, desc: InputDataDescriptor, i: int, ko, device):
        ss, ts = desc.stack_structure, desc.tensors_structure
        sp0 = int(ss.number_of_points_per_stack[:i].sum())
        sp1 = sp0 + int(ss.number_of_points_per_stack[i])
        or0 = int(ss.number_of_orientations_per_stack[:i].sum())
        or1 = or0 + int(ss.number_of_orientations_per_stack[i])
        su0 = int(ss.number_of_surfaces_per_stack[:i].sum())
        su1 = su0 + int(ss.number_of_surfaces_per_stack[i])
        self.sp_slice = slice(sp0, sp1)
        self.surf_slice = slice(su0, su1)
        nps = np.asarray(ts.number_of_points_per_surface[su0:su1], dtype=np.int64)
        if (nps < 1).any():
            raise ValueError(f"stack {i}: every surface needs at least one surface point")
        sp = ii.surface_points.sp_coords[sp0:sp1]
        nug = ii.surface_points.nugget_effect_scalar[sp0:sp1]
        starts = np.concatenate([[0], np.cumsum(nps)[:-1]]).astype(np.int64)
        is_ref = np.zeros(sp.shape[0], bool)
        is_ref[starts] = True
        reps = nps - 1
        self.ref_local = starts                                   # reference point of each surface, stack-local index
        self.is_ref = is_ref
        self.n_surf = int(nps.shape[0])
        self.n_rest = int(reps.sum())
        self.n_ori = or1 - or0
        self.n_drift = _n_drift(ko.uni_degree)
        rest = sp[~is_ref]
        ref = np.repeat(sp[starts], reps, axis=0)
        row_nug = 0.5 * (nug[~is_ref] + np.repeat(nug[starts], reps))
        surf_off = np.concatenate([[0], np.cumsum(reps)]).astype(np.int32)
        dev = lambda a, dt=F64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)
        self.rest = dev(rest.T)
        self.ref = dev(ref.T)
        self.sp_nugget = dev(row_nug)
        self.ref_unique = dev(sp[starts].T)
        self.surf_offsets = dev(surf_off, torch.int32)
        self.ori_pos = dev(ii.orientations.dip_positions[or0:or1].T)
        self.ori_grad = dev(ii.orientations.dip_gradients[or0:or1].T)
        self.ori_nugget = dev(ii.orientations.nugget_effect_grad[or0:or1])
        # gather indices of rest / ref points inside the stack's surface-point table (for the fault tables)
        self.rest_idx = torch.as_tensor(np.nonzero(~is_ref)[0], device=device)
        self.ref_idx = torch.as_tensor(np.repeat(starts, reps), device=device)
        self.ref_local_dev = torch.as_tensor(starts, device=device)
        fr = ss.faults_relations
        active = np.nonzero(np.asarray(fr)[:, i])[0] if fr is not None else np.zeros(0, dtype=np.int64)
        self.active_faults = active
        self.active_faults_dev = torch.as_tensor(active, device=device) if active.size else None
        self.ko = ko
        self.fault_rest = None
        self.fault_ref = None
        self.n_faults = 0
        self._struct = None                                        # cached ctypes view of the tables (see struct())

    def set_faults(self, fault_on_sp: Optional[torch.Tensor]):
        """fault_on_sp: [n_f, n_sp_of_stack] values of the active fault blocks at this stack's surface points."""
        self._struct = None
        if fault_on_sp is None or fault_on_sp.shape[0] == 0:
            self.fault_rest = self.fault_ref = None
            self.n_faults = 0
            return
        self.n_faults = int(fault_on_sp.shape[0])
        self.fault_rest = fault_on_sp.index_select(1, self.rest_idx).contiguous()
        self.fault_ref = fault_on_sp.index_select(1, self.ref_idx).contiguous()

    def struct(self) -> _lib.GpbStack:
        if self._struct is None:
            self._struct = self._make_struct()
        return self._struct

    def _make_struct(self) -> _lib.GpbStack:
        ko = self.ko
        return _lib.GpbStack(
            self.n_ori, self.n_rest, self.n_surf, self.n_drift, self.n_faults, _kernel_code(ko.kernel_function),
            float(ko.range), float(ko.c_o), float(ko.i_res), float(ko.gi_res),
            _ptr(self.ori_pos), _ptr(self.ori_grad), _ptr(self.ori_nugget), _ptr(self.rest), _ptr(self.ref),
            _ptr(self.sp_nugget), _ptr(self.fault_rest), _ptr(self.fault_ref), _ptr(self.surf_offsets),
            _ptr(self.ref_unique))

    @property
    def n(self) -> int:
        return 3 * self.n_ori + self.n_rest + self.n_drift + self.n_faults


class ModelTables(list):
    """Per-stack device tables plus the model-wide constants (all surface points, unit ids), uploaded once per
    ``compute_model`` call: no host-to-device copy is issued inside the per-level loop, so the host runs ahead of the
    GPU instead of synchronising on every small pageable copy."""

    def __init__(self, ii: InterpolationInput, desc: InputDataDescriptor, ko, device):
        super().__init__(StackTables(ii, desc, i, ko, device) for i in range(desc.stack_structure.n_stacks))
        self.sp_all = torch.as_tensor(np.ascontiguousarray(ii.surface_points.sp_coords.T), dtype=F64, device=device)
        self.unit_values = torch.as_tensor(np.asarray(ii.unit_values, dtype=np.float64), device=device)
        self.src_cache = {}            # stack -> packed evaluation table (the weights do not change between levels)


# ------------------------------------------------------------------------------------------------ engine
@dataclass
class FieldsOnDevice:
    """Result of evaluating all stacks on one domain (device tensors)."""
    segments: List[Segment]
    grid_size: int
    Z: torch.Tensor                    # [n_stacks, L]   L = grid_size + n_sp
    G: Optional[torch.Tensor]          # [n_stacks, 3, L] or None
    block: torch.Tensor                # [n_stacks, L]
    final_block: torch.Tensor          # [L]
    faults_block: torch.Tensor         # [L]
    squeezed: torch.Tensor             # [n_stacks, L] uint8
    mask: torch.Tensor                 # [n_stacks, L] uint8
    isovalues: List[torch.Tensor]      # per stack [n_surf]
    weights: List[torch.Tensor]
    cond: List[Optional[float]]
    srcs: List[torch.Tensor] = None    # packed evaluation tables per stack

    def seg_slice(self, name: str) -> slice:
        off = 0
        for s in self.segments:
            if s.name == name:
                return slice(off, off + s.m)
            off += s.m
        return slice(0, 0)


class B200Engine:
    def __init__(self, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise _lib.GpbError("the B200 backend needs a CUDA device (no CPU fallback)")
        self.lib = _lib.lib()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self.lib.gpb_device_info(self.device_index, C.byref(sm), C.byref(ma), C.byref(mi)))
        self.sm_count = sm.value
        self._held_stream = None

    # -- helpers --------------------------------------------------------------------------------------------
    @property
    def stream(self) -> int:
        """The caller's current CUDA stream.  Inside ``hold_stream()`` the lookup is done once (786 lookups cost 6 ms of
        the 100 ms a 15-stack octree-8 ``compute_model`` takes on the host)."""
        if self._held_stream is not None:
            return self._held_stream
        return torch.cuda.current_stream(self.device).cuda_stream

    @contextlib.contextmanager
    def hold_stream(self):
        outer = self._held_stream
        if outer is None:
            self._held_stream = torch.cuda.current_stream(self.device).cuda_stream
        try:
            yield
        finally:
            self._held_stream = outer

    def empty(self, *shape, dtype=F64) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # -- stages ----------------------------------------------------------------------------------------------
    def assemble(self, st: StackTables, extra_rows: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """System matrix (column-major: tensor row j = matrix column j) and right-hand side.  With ``extra_rows`` the
        leading dimension is n + extra_rows rounded up to an even number (16-byte aligned columns; the symmetric solve
        carries its right-hand sides as extra rows) and the tensor has shape (n, lda)."""
        n = st.n
        lda = n if extra_rows == 0 else (n + extra_rows + 1) & ~1
        A = self.empty(n, lda)              # symmetric at this point
        b = self.empty(n)
        s = st.struct()
        _lib.check(self.lib.gpb_assemble_cov(C.byref(s), _ptr(A), lda, _ptr(b), self.stream))
        return A, b

    def _check_info(self, info: torch.Tensor, what: str) -> int:
        k = int(info.item())                # one 4-byte read per solve: a singular system must not pass silently
        if k != 0:
            raise _lib.GpbError(f"{what}: zero pivot at column {k} -- the co-kriging system is singular "
                                "(duplicate surface points / orientations with zero nugget, or an all-zero drift column)")
        return k

    def solve(self, A: torch.Tensor, b: torch.Tensor, what: str = "LU solve") -> torch.Tensor:
        """General path: b -> A^-1 b by the pivoted blocked LU (A is overwritten by its factors); raises on a zero
        pivot.  A large system of odd order is first copied into a buffer with an even leading dimension (every column
        then starts on a 16-byte boundary and the K = 256 trailing update uses 16-byte operand copies)."""
        n = A.shape[0]
        lda = A.shape[1]
        if lda == n and n % 2 == 1 and n >= int(self.lib.gpb_lu_set_outer_min_n(-1)):
            W = self.empty(n, n + 1)
            W[:, :n].copy_(A)
            A, lda = W, n + 1
        ipiv = self.empty(n, dtype=torch.int32)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.gpb_lu_solve(n, _ptr(A), lda, _ptr(b), 1, n, _ptr(ipiv), _ptr(info), self.stream))
        self._check_info(info, what)
        return b

    SMALL_N = 160          # systems up to this order are solved by the one-CTA LU (every reference example model)

    def solve_stack(self, st: StackTables, what: str = "stack") -> Tuple[torch.Tensor, str]:
        """Assemble and solve one stack's saddle-point system; returns (weights, path).  Systems larger than SMALL_N
        take the symmetric path (Cholesky of the covariance block + Schur complement of the drift rows,
        gpb_sym_solve); if the covariance block is not numerically positive definite the system is re-assembled and
        solved by the pivoted LU."""
        n = st.n
        nk = 3 * st.n_ori + st.n_rest
        if n > self.SMALL_N and nk >= 1 and os.environ.get("GPB_SOLVER", "sym") != "lu":
            A, b = self.assemble(st, extra_rows=1)
            info = torch.zeros(1, dtype=torch.int32, device=self.device)
            _lib.check(self.lib.gpb_sym_solve(n, nk, _ptr(A), A.shape[1], _ptr(b), 1, n, _ptr(info), self.stream))
            if int(info.item()) == 0:
                return b, "sym"
            del A, b
        A, b = self.assemble(st)
        return self.solve(A, b, what), "lu"

    def pack(self, st: StackTables, w: torch.Tensor) -> torch.Tensor:
        s = st.struct()
        nd = int(self.lib.gpb_eval_table_doubles(C.byref(s)))
        src = self.empty(nd)
        _lib.check(self.lib.gpb_pack_eval_table(C.byref(s), _ptr(w), _ptr(src), self.stream))
        return src

    def evaluate_segment(self, st: StackTables, src: torch.Tensor, seg: Segment, off: int, Z: torch.Tensor,
                         G: Optional[torch.Tensor], fault_vals: Optional[torch.Tensor]):
        """Z: [L] row of this stack; G: [3, L] or None; fault_vals: [n_f, L] or None; off = column offset."""
        if seg.m == 0:
            return
        s = st.struct()
        L = Z.shape[0]
        o8 = off * 8
        gp = [None, None, None] if G is None else [_ptr(G[a], o8) for a in range(3)]
        fv = _ptr(fault_vals, o8) if fault_vals is not None else None
        if seg.grid is not None:
            _lib.check(self.lib.gpb_eval_regular(C.byref(s), _ptr(src), C.byref(seg.grid), seg.i0, seg.i0 + seg.m,
                                                 fv, L, _ptr(Z, o8), gp[0], gp[1], gp[2], self.stream))
        else:
            _lib.check(self.lib.gpb_eval_points(C.byref(s), _ptr(src), _ptr(seg.xyz), seg.xyz.shape[1], seg.m,
                                                fv, L, _ptr(Z, o8), gp[0], gp[1], gp[2], self.stream))

    # -- all stacks on one domain ----------------------------------------------------------------------------
    def interpolate_all_fields(self, ii: InterpolationInput, options: InterpolationOptions, desc: InputDataDescriptor,
                               segments: List[Segment], weights_cache: List[Optional[torch.Tensor]],
                               gradient: Optional[bool] = None, tables: Optional[List[StackTables]] = None,
                               comm: Optional[Comm] = None) -> FieldsOnDevice:
        """`segments` are this rank's evaluation points.  With a multi-rank `comm`, the systems are solved once (rank 0,
        or the owner rank for fault-free stacks) and the weights broadcast, and the fault-block minima are all-reduced,
        so every rank sees the values a single-GPU run would produce."""
        comm = comm or Comm()
        ko = options.kernel_options
        if gradient is None:
            gradient = bool(options.evaluation_options.compute_scalar_gradient)
        ss = desc.stack_structure
        n_st = ss.n_stacks
        rel = [_rel_code(r) for r in ss.masking_descriptor]
        if tables is None or not isinstance(tables, ModelTables):
            tables = ModelTables(ii, desc, ko, self.device)
        sp_all = tables.sp_all
        n_sp = sp_all.shape[1]
        gsz = sum(s.m for s in segments)
        L = gsz + n_sp
        sp_seg = Segment("surface_points", n_sp, xyz=sp_all)
        Z = self.empty(n_st, L)
        G = self.empty(n_st, 3, L) if gradient else None
        block = self.empty(n_st, L)
        values_everywhere = self.empty(n_st, L)
        unit_values = tables.unit_values
        iso_min = self.empty(n_st)
        iso_max = self.empty(n_st)
        isos, conds, srcs = [], [], []
        tmp_min = self.empty(1)
        # Stacks whose system has no fault-drift column depend on nothing: with several ranks their assemble + solve are
        # dealt out round-robin and every owner broadcasts its weights (sharding by independent stack / series,
        # SURVEY.md 8e); fault-dependent stacks are solved on rank 0 in stack order below.
        pre_cond = {}
        free = [i for i in range(n_st) if weights_cache[i] is None and tables[i].active_faults_dev is None]
        if comm.world > 1 and len(free) > 1:
            want_cond = bool(getattr(ko, "compute_condition_number", False))
            mine = {}
            for k, i in enumerate(free):
                if k % comm.world == comm.rank:
                    tables[i].set_faults(None)
                    A, b = self.assemble(tables[i])
                    c = float(torch.linalg.cond(A).item()) if want_cond else float("nan")
                    mine[i] = (self.solve(A, b), c)
                    del A
            for k, i in enumerate(free):
                owner = k % comm.world
                w_new, c = mine[i] if owner == comm.rank else (self.empty(tables[i].n), float("nan"))
                weights_cache[i] = comm.broadcast(w_new, src=owner)
                if want_cond:
                    ct = comm.broadcast(torch.tensor([c], dtype=F64, device=self.device), src=owner)
                    pre_cond[i] = float(ct.item())
        for i in range(n_st):
            st = tables[i]
            f_every = None
            if st.active_faults_dev is not None:
                f_every = values_everywhere.index_select(0, st.active_faults_dev).contiguous()
                st.set_faults(f_every[:, gsz:][:, st.sp_slice])
            else:
                st.set_faults(None)
            cond = pre_cond.get(i)
            if weights_cache[i] is None:
                if comm.rank == 0:
                    A, b = self.assemble(st)
                    if getattr(ko, "compute_condition_number", False):
                        cond = float(torch.linalg.cond(A).item())
                    w_new = self.solve(A, b)
                    del A
                else:
                    w_new = self.empty(st.n)
                weights_cache[i] = comm.broadcast(w_new, src=0)
            w = weights_cache[i]
            if w.shape[0] != st.n:
                raise ValueError(f"stack {i}: cached weights have length {w.shape[0]}, system size is {st.n}")
            src = tables.src_cache.get(i)
            if src is None:
                src = tables.src_cache[i] = self.pack(st, w)
            srcs.append(src)
            Gi = G[i] if gradient else None
            # surface points first (the isovalues feed the activator), then every grid segment
            self.evaluate_segment(st, src, sp_seg, gsz, Z[i], Gi, f_every)
            off = 0
            for seg in segments:
                self.evaluate_segment(st, src, seg, off, Z[i], Gi, f_every)
                off += seg.m
            iso = Z[i, gsz + st.sp_slice.start:gsz + st.sp_slice.stop].index_select(0, st.ref_local_dev).contiguous()
            isos.append(iso)
            iso_min[i] = iso.min()
            iso_max[i] = iso.max()
            ids = unit_values[st.surf_slice.start:st.surf_slice.stop + 1].contiguous()
            if ids.shape[0] != st.n_surf + 1:
                raise ValueError("unit_values must hold one id per surface plus the basement")
            _lib.check(self.lib.gpb_activate(_ptr(Z[i]), L, _ptr(iso), _ptr(ids), st.n_surf, float(options.sigmoid_slope),
                                             _ptr(block[i]), self.stream))
            if rel[i] == StackRelationType.FAULT.value:
                _lib.check(self.lib.gpb_min(_ptr(block[i]), L, _ptr(tmp_min), self.stream))
                comm.all_reduce_min(tmp_min)
                _lib.check(self.lib.gpb_shift(_ptr(block[i]), L, _ptr(tmp_min), _ptr(values_everywhere[i]), self.stream))
            else:
                values_everywhere[i].copy_(block[i])
            conds.append(cond)
            if cond is not None:
                ko.condition_number = cond
        final_block = self.empty(L)
        faults_block = self.empty(L)
        squeezed = self.empty(n_st, L, dtype=torch.uint8)
        mask = self.empty(n_st, L, dtype=torch.uint8)
        rel_arr = (C.c_int * n_st)(*rel)
        _lib.check(self.lib.gpb_combine(_ptr(Z), _ptr(block), L, L, n_st, rel_arr, _ptr(iso_min), _ptr(iso_max),
                                        _ptr(final_block), _ptr(faults_block), _ptr(squeezed), _ptr(mask), self.stream))
        return FieldsOnDevice(segments, gsz, Z, G, block, final_block, faults_block, squeezed, mask, isos,
                              [weights_cache[i] for i in range(n_st)], conds, srcs)

    def gradient_at(self, st: StackTables, src: torch.Tensor, xyz: torch.Tensor) -> torch.Tensor:
        """Engine-convention gradient [3, m] of one stack's field at explicit points.  The fault drift has no
        gradient term, so the fault columns are skipped."""
        m = xyz.shape[1]
        Z = self.empty(m)
        G = self.empty(3, m)
        s = st.struct()
        s.n_faults = 0
        _lib.check(self.lib.gpb_eval_points(C.byref(s), _ptr(src), _ptr(xyz), m, m, None, 0, _ptr(Z), _ptr(G[0]), _ptr(G[1]),
                                            _ptr(G[2]), self.stream))
        return G

    # -- octree ----------------------------------------------------------------------------------------------
    def corners_of(self, centers: torch.Tensor, d: np.ndarray) -> torch.Tensor:
        nv = centers.shape[1]
        out = self.empty(3, 8 * nv)
        _lib.check(self.lib.gpb_voxel_corners(_ptr(centers), nv, nv, d[0] / 2, d[1] / 2, d[2] / 2, _ptr(out), 8 * nv,
                                              self.stream))
        return out

    def mark(self, nv: int, lith_corners: torch.Tensor, fault_corners: torch.Tensor, force_all: bool) -> torch.Tensor:
        """Refinement test on nv voxels given the ids at their 8 corners each (uint8 marks)."""
        mark = self.empty(nv, dtype=torch.uint8)
        if nv:
            _lib.check(self.lib.gpb_mark_voxels(_ptr(lith_corners), _ptr(fault_corners), nv, int(force_all), _ptr(mark),
                                                self.stream))
        return mark

    def emit(self, centers: torch.Tensor, d: np.ndarray, mark: torch.Tensor) -> torch.Tensor:
        """Children (8 per marked voxel, parent order preserved) of the marked voxels."""
        nv = centers.shape[1]
        n_children = C.c_longlong(0)
        _lib.check(self.lib.gpb_emit_children(_ptr(centers), nv, nv, _ptr(mark), d[0] / 4, d[1] / 4, d[2] / 4, None, 0,
                                              C.byref(n_children), self.stream))
        nc = int(n_children.value)
        children = self.empty(3, nc)
        if nc:
            _lib.check(self.lib.gpb_emit_children(_ptr(centers), nv, nv, _ptr(mark), d[0] / 4, d[1] / 4, d[2] / 4,
                                                  _ptr(children), nc, C.byref(n_children), self.stream))
        return children

    def gather_fields(self, f: FieldsOnDevice, totals: Sequence[int], comm: Comm) -> FieldsOnDevice:
        """All-gather a range-sharded level into whole arrays (segment by segment; the surface-point tail is
        replicated).  Identity on a single rank."""
        if comm.world == 1:
            return f
        n_sp = f.Z.shape[1] - f.grid_size

        def full(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
            if t is None:
                return None
            parts, off = [], 0
            for seg, tot in zip(f.segments, totals):
                parts.append(comm.all_gather_cat(t[..., off:off + seg.m].contiguous(), tot))
                off += seg.m
            parts.append(t[..., off:off + n_sp])
            return torch.cat(parts, dim=-1).contiguous()

        segs = [Segment(sg.name, int(tot)) for sg, tot in zip(f.segments, totals)]
        return FieldsOnDevice(segs, int(sum(totals)), full(f.Z), full(f.G), full(f.block), full(f.final_block),
                              full(f.faults_block), full(f.squeezed), full(f.mask), f.isovalues, f.weights, f.cond, f.srcs)


# ------------------------------------------------------------------------------------------------ dense field API
def compute_dense_fields(interpolation_input: InterpolationInput, options: InterpolationOptions,
                         data_descriptor: InputDataDescriptor, *, stack: int = 0, engine: Optional[B200Engine] = None,
                         point_range: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None,
                         n_slabs: int = 8, device: Optional[int] = None) -> torch.Tensor:
    """Scalar field and gradient of one (fault-free) stack on the dense regular grid, host in / host out.

    The reference obtains these with ``compute_model`` + ``evaluation_options.compute_scalar_gradient = True`` and reads
    ``exported_fields_dense_grid.{scalar_field, gx_field, gy_field, gz_field}``.  This is the streaming form of that
    call for grids whose outputs are too large to keep (512^3: 4.3 GB): tables H2D, assemble, solve, then the grid
    range is evaluated slab by slab on the compute stream while a copy stream drains finished slabs into pinned host
    memory.  Returns a pinned ``[4, m]`` tensor: Z, gx, gy, gz of points [i0, i1)."""
    eng = engine or B200Engine(device)
    ii, desc, ko = interpolation_input, data_descriptor, options.kernel_options
    g = ii.grid.dense_grid
    if g is None:
        raise ValueError("compute_dense_fields needs a dense grid")
    i0, i1 = point_range if point_range is not None else (0, g.n_points)
    m = i1 - i0
    st = StackTables(ii, desc, stack, ko, eng.device)
    fr = desc.stack_structure.faults_relations
    if fr is not None and np.asarray(fr)[:, stack].any():
        raise ValueError("compute_dense_fields handles fault-free stacks; use compute_model for faulted ones")
    A, b = eng.assemble(st)
    w = eng.solve(A, b)
    del A
    src = eng.pack(st, w)
    if out is None:
        out = torch.empty((4, m), dtype=F64, pin_memory=True)
    gd = regular_descriptor(g)
    nyz = int(g.regular_grid_shape[1] * g.regular_grid_shape[2])
    # slabs aligned to whole x planes when possible (keeps every slab on the z-run kernel)
    per = max(1, -(-m // max(1, n_slabs)))
    if per > nyz:
        per = -(-per // nyz) * nyz
    compute = torch.cuda.current_stream(eng.device)
    copier = torch.cuda.Stream(eng.device)
    bufs = [eng.empty(4, per) for _ in range(2)]
    free_ev = [None, None]
    k = 0
    for s0 in range(0, m, per):
        s1 = min(m, s0 + per)
        buf = bufs[k & 1]
        if free_ev[k & 1] is not None:
            compute.wait_event(free_ev[k & 1])          # the copy that used this buffer has finished
        seg = Segment("dense_grid", s1 - s0, grid=gd, i0=i0 + s0)
        eng.evaluate_segment(st, src, seg, 0, buf[0], buf[1:], None)
        done = torch.cuda.Event()
        done.record(compute)
        with torch.cuda.stream(copier):
            copier.wait_event(done)
            for a in range(4):                          # row by row: contiguous 1-D copies stay on the DMA path
                out[a, s0:s1].copy_(buf[a, :s1 - s0], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copier)
            free_ev[k & 1] = ev
        k += 1
    copier.synchronize()
    return out


# ------------------------------------------------------------------------------------------------ materialisation
def _np(t: Optional[torch.Tensor]) -> Optional[np.ndarray]:
    return None if t is None else t.detach().cpu().numpy()


def _level_outputs(f: FieldsOnDevice, grid: EngineGrid, rel_enum: Sequence) -> List[InterpOutput]:
    """Host containers over the device results; every array is copied on first access only."""
    D = Deferred
    fb = D(lambda: _np(f.final_block))
    fa = D(lambda: _np(f.faults_block))
    outs = []
    for i in range(f.Z.shape[0]):
        g = (None, None, None) if f.G is None else tuple(D(lambda i=i, a=a: _np(f.G[i, a])) for a in range(3))
        ef = ExportedFields(D(lambda i=i: _np(f.Z[i])), g[0], g[1], g[2], f.grid_size, D(lambda i=i: _np(f.isovalues[i])))
        sfo = ScalarFieldOutput(D(lambda i=i: _np(f.weights[i])), grid, ef, D(lambda i=i: _np(f.block[i])[None, :]),
                                rel_enum[i], D(lambda i=i: _np(f.mask[i]).astype(bool)))
        comb = CombinedScalarFieldsOutput(D(lambda i=i: _np(f.squeezed[i]).astype(bool)), fb, fa)
        outs.append(InterpOutput(sfo, comb))
    return outs


def _fill_regular_from_octree(levels_host, base_shape: np.ndarray, key) -> np.ndarray:
    """Dense array at the finest octree resolution: level-0 values upsampled, refined voxels overwritten by their
    children (the engine's octree -> regular fill used by RawArraysSolution)."""
    shape = np.asarray(base_shape, dtype=int)
    vals = np.asarray(key(levels_host[0])).reshape(shape)
    # index arrays of the voxels of each level inside the level's full lattice
    ijk = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), axis=-1).reshape(-1, 3)
    dense = vals
    for lvl in range(1, len(levels_host)):
        sel = levels_host[lvl - 1]["selected"]
        sel = sel.get() if isinstance(sel, Deferred) else sel
        dense = dense.repeat(2, axis=0).repeat(2, axis=1).repeat(2, axis=2)
        parents = ijk[sel]
        off = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)])
        ijk = (parents[:, None, :] * 2 + off[None, :, :]).reshape(-1, 3)
        v = key(levels_host[lvl])
        dense[ijk[:, 0], ijk[:, 1], ijk[:, 2]] = v
    return dense.ravel()


# ------------------------------------------------------------------------------------------------ triangulation
def triangulate(valid: np.ndarray, ijk: np.ndarray) -> np.ndarray:
    """Quads (as two triangles) around every crossed edge shared by four existing surface voxels; vertex index =
    rank among voxels with at least one crossing (same rule as oracle.dual_contour_triangles, vectorised)."""
    vv = valid.any(axis=1)
    k = ijk[vv].astype(np.int64)
    val = valid[vv]
    if k.shape[0] == 0:
        return np.zeros((0, 3), dtype=np.int64)
    lo = k.min(axis=0) - 1
    span = k.max(axis=0) - lo + 2
    code = lambda a: ((a[:, 0] - lo[0]) * span[1] + (a[:, 1] - lo[1])) * span[2] + (a[:, 2] - lo[2])
    codes = code(k)
    order = np.argsort(codes, kind="stable")
    sorted_codes = codes[order]

    def find(a):
        c = code(a)
        pos = np.searchsorted(sorted_codes, c)
        pos = np.clip(pos, 0, len(sorted_codes) - 1)
        ok = sorted_codes[pos] == c
        return np.where(ok, order[pos], -1)

    tris = []
    hh_edge = {0: 3, 1: 7, 2: 11}
    others = {0: (1, 2), 1: (0, 2), 2: (0, 1)}
    for ax in range(3):
        e = hh_edge[ax]
        u, v = others[ax]
        n = np.nonzero(val[:, e])[0]
        if n.size == 0:
            continue
        ku = k[n].copy(); ku[:, u] += 1
        kv = k[n].copy(); kv[:, v] += 1
        kuv = ku.copy(); kuv[:, v] += 1
        a, b, c = find(ku), find(kv), find(kuv)
        ok = (a >= 0) & (b >= 0) & (c >= 0)
        n, a, b, c = n[ok], a[ok], b[ok], c[ok]
        t = np.stack([np.stack([n, a, c], axis=1), np.stack([n, c, b], axis=1)], axis=1).reshape(-1, 3)
        tris.append(t)
    return np.concatenate(tris).astype(np.int64) if tris else np.zeros((0, 3), dtype=np.int64)


# ------------------------------------------------------------------------------------------------ entry point
def compute_model(interpolation_input: InterpolationInput, options: InterpolationOptions,
                  data_descriptor: InputDataDescriptor, geophysics_input=None, *, device: Optional[int] = None,
                  engine: Optional[B200Engine] = None, comm: Optional[Comm] = None) -> Solutions:
    """Drop-in for ``gempy_engine.compute_model`` (same positional/keyword signature; the keyword-only extras
    select the CUDA device and, for one-process-per-GPU runs, the torch.distributed group).  Every rank returns
    the same, complete ``Solutions``.  Raises ``NotImplementedError`` for geophysics input."""
    comm = comm or Comm()
    if geophysics_input is not None and interpolation_input.grid.geophysics_grid is None:
        raise ValueError("geophysics_input needs a centered (geophysics) grid")
    if geophysics_input is not None and getattr(geophysics_input, "magnetics_input", None) is not None:
        raise NotImplementedError("magnetics is outside the B200 backend's scope")
    eng = engine or B200Engine(device)
    with eng.hold_stream():
        return _compute_model(eng, interpolation_input, options, data_descriptor, geophysics_input, comm)


def _compute_model(eng: B200Engine, interpolation_input, options, data_descriptor, geophysics_input, comm: Comm) -> Solutions:
    ii, desc = interpolation_input, data_descriptor
    eo = options.evaluation_options
    grid = ii.grid
    ss = desc.stack_structure
    n_st = ss.n_stacks
    rel_enum = list(ss.masking_descriptor)
    ko = options.kernel_options
    if grid.octree_grid is None:
        raise ValueError("the engine grid always carries an octree grid (_engine_factory.py:88-96)")

    cache: List[Optional[torch.Tensor]] = [None] * n_st
    if ii.weights:
        for i, w in enumerate(ii.weights):
            if w is not None and len(w):
                cache[i] = torch.as_tensor(np.asarray(w, dtype=np.float64), device=eng.device)
    tables = ModelTables(ii, desc, ko, eng.device)

    n_levels = int(eo.number_octree_levels)
    dc_level = min(int(eo.number_octree_levels_surface), n_levels) - 1 if eo.mesh_extraction else -1
    centers, d = regular_centers_device(grid.octree_grid, eng.device)
    extra_full: List[Segment] = []
    if grid.dense_grid is not None:
        extra_full.append(Segment("dense_grid", grid.dense_grid.n_points, grid=regular_descriptor(grid.dense_grid)))
    for name in ("custom_grid", "topography", "sections", "geophysics_grid"):
        g = getattr(grid, name)
        if g is not None:
            extra_full.append(Segment(name, g.n_points,
                                      xyz=torch.as_tensor(np.ascontiguousarray(g.values.T), dtype=F64, device=eng.device)))

    def local_part(seg: Segment) -> Segment:
        """This rank's contiguous share of a segment."""
        i0, i1 = comm.shard(seg.m)
        if seg.grid is not None:
            return Segment(seg.name, i1 - i0, grid=seg.grid, i0=seg.i0 + i0)
        return Segment(seg.name, i1 - i0, xyz=seg.xyz[:, i0:i1].contiguous())

    octree_levels: List[OctreeLevel] = []
    levels_host = []
    prev_regular = grid.octree_grid
    dc_payload = None
    gravity = None
    for lvl in range(n_levels):
        need_corners = (lvl < n_levels - 1) or (lvl == dc_level)
        nv = centers.shape[1]
        v0, v1 = comm.shard(nv)
        centers_loc = centers[:, v0:v1].contiguous()
        segs = [Segment("octree_grid", v1 - v0, xyz=centers_loc)]
        totals = [nv]
        if lvl == 0:
            for sg in extra_full:
                segs.append(local_part(sg))
                totals.append(sg.m)
        corners = None
        if need_corners:
            corners_loc = eng.corners_of(centers_loc, d)
            segs.append(Segment("corners", corners_loc.shape[1], xyz=corners_loc))
            totals.append(8 * nv)
            corners = corners_loc if comm.world == 1 else eng.corners_of(centers, d)
        f_loc = eng.interpolate_all_fields(ii, options, desc, segs, cache, tables=tables, comm=comm)
        # ---- refinement marks: local test, all-gathered so that every rank emits the identical child list
        mark_full = None
        if lvl < n_levels - 1:
            csl = f_loc.seg_slice("corners")
            mark_loc = eng.mark(v1 - v0, f_loc.final_block[csl].contiguous(), f_loc.faults_block[csl].contiguous(),
                                force_all=lvl < int(eo.octree_min_level))
            mark_full = comm.all_gather_cat(mark_loc, nv)
        f = eng.gather_fields(f_loc, totals, comm)
        # ---- host containers of this level
        centers_host = Deferred(lambda c=centers: _np(c).T.copy())     # explicit centres (shift included), as evaluated
        if lvl == 0:
            og0 = RegularGrid(grid.octree_grid.orthogonal_extent, grid.octree_grid.regular_grid_shape)
            og0._values, og0._n_points = centers_host, nv
            lvl_grid = EngineGrid(octree_grid=og0, dense_grid=grid.dense_grid, topography=grid.topography,
                                  sections=grid.sections, custom_grid=grid.custom_grid,
                                  geophysics_grid=grid.geophysics_grid)
        else:
            og = RegularGrid.from_octree_level(centers_host, prev_regular)
            og._dxdydz, og._n_points = d.copy(), nv
            prev_regular = og
            lvl_grid = EngineGrid(octree_grid=og)
        outs = _level_outputs(f, lvl_grid, rel_enum)
        level = OctreeLevel(grid_centers=lvl_grid, outputs_centers=outs,
                            _grid_corners=None if corners is None else
                            Deferred(lambda c=corners: EngineGrid.from_xyz_coords(_np(c).T)))
        level._device_fields = f
        octree_levels.append(level)
        host = {"lith": Deferred(lambda o=outs[-1], nv=nv: np.rint(o.combined_scalar_field.final_block[:nv])),
                "faults": Deferred(lambda o=outs[-1], nv=nv: np.rint(o.combined_scalar_field.faults_block[:nv])),
                "selected": None}
        levels_host.append(host)
        if lvl == 0 and geophysics_input is not None:
            gravity = _forward_gravity(eng, geophysics_input, grid.geophysics_grid, f)
        if lvl == dc_level:
            dc_payload = (centers, d.copy(), corners, f)
        if lvl == n_levels - 1:
            break
        level._marked_voxels = Deferred(lambda mk=mark_full: _np(mk).astype(bool))
        host["selected"] = level._marked_voxels
        centers = eng.emit(centers, d, mark_full)
        d = d / 2

    meshes = None
    if eo.mesh_extraction and dc_payload is not None:
        meshes = _dual_contouring(eng, ii, options, desc, tables, cache, dc_payload, grid.octree_grid)

    sol = Solutions(octree_levels, meshes, gravity, options.block_solutions_type)
    sol.raw_arrays = _raw_arrays(sol, levels_host, grid, options, meshes)
    return sol


def _forward_gravity(eng: B200Engine, geophysics_input, centered_grid, f: FieldsOnDevice) -> np.ndarray:
    """gravity[c] = sum_k tz[k] * density[lith id at (centre c, kernel voxel k)] on the (gathered) level-0 fields."""
    tz = getattr(geophysics_input, "tz", None)
    dens = getattr(geophysics_input, "densities", None)
    if tz is None or dens is None:
        gi = getattr(geophysics_input, "gravity_input", None)
        tz, dens = gi.tz, gi.densities
    tz_d = torch.as_tensor(np.ascontiguousarray(tz, dtype=np.float64), device=eng.device)
    dens_d = torch.as_tensor(np.ascontiguousarray(dens, dtype=np.float64), device=eng.device)
    n_centers = int(centered_grid.centers.shape[0])
    n_k = int(centered_grid.kernel_grid_centers.shape[0])
    if tz_d.shape[0] != n_k:
        raise ValueError(f"tz has {tz_d.shape[0]} entries, the centered grid kernel has {n_k} voxels")
    sl = f.seg_slice("geophysics_grid")
    block = f.final_block[sl].contiguous()
    out = eng.empty(n_centers)
    _lib.check(eng.lib.gpb_gravity(_ptr(block), _ptr(dens_d), int(dens_d.shape[0]), _ptr(tz_d), n_centers, n_k, _ptr(out),
                                   eng.stream))
    return _np(out)


def _dual_contouring(eng: B200Engine, ii, options, desc, tables, cache, payload, root_grid) -> List[DualContouringMesh]:
    centers, d, corners, f = payload
    nv = centers.shape[1]
    csl = f.seg_slice("corners")
    e = root_grid.orthogonal_extent
    ijk = np.rint((_np(centers).T - GRID_SHIFT - e[[0, 2, 4]]) / d - 0.5).astype(np.int64)
    ss = desc.stack_structure
    rel = [_rel_code(r) for r in ss.masking_descriptor]
    lib = eng.lib
    # pass 1: every surface's device work is queued without a host round trip; the results start their way to the host
    # on the same stream (non-blocking copies).  pass 2 triangulates on the host while later surfaces still run.
    pending = []
    for i in range(ss.n_stacks):
        Zc = f.Z[i, csl].contiguous()
        if rel[i] == StackRelationType.FAULT.value:
            own = None
        else:
            own = eng.empty(nv, dtype=torch.uint8)
            sq = f.squeezed[i, csl].contiguous()
            _lib.check(lib.gpb_any8(_ptr(sq), nv, _ptr(own), eng.stream))
        iso_host = _np(f.isovalues[i])
        for s_idx, iso in enumerate(iso_host):
            valid = eng.empty(12 * nv, dtype=torch.uint8)
            xyz_e = eng.empty(3, 12 * nv)
            _lib.check(lib.gpb_dc_edges(_ptr(corners), 8 * nv, _ptr(Zc), nv, float(iso), _ptr(own), _ptr(valid),
                                        _ptr(xyz_e), eng.stream))
            # gradient of stack i's field at the crossings: one more fused evaluation on the edge points
            grad = eng.gradient_at(tables[i], f.srcs[i], xyz_e)
            verts = eng.empty(3, nv)
            _lib.check(lib.gpb_dc_vertices(_ptr(valid), _ptr(xyz_e), _ptr(grad), nv, 1.0, _ptr(verts), eng.stream))
            host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (valid, verts, xyz_e, grad)]
            for h, t in zip(host, (valid, verts, xyz_e, grad)):
                h.copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(eng.device))
            pending.append((host, done, (valid, verts, xyz_e, grad)))       # device tensors stay alive until copied
    meshes: List[DualContouringMesh] = []
    for host, done, _keep in pending:
        done.synchronize()
        valid_h = host[0].numpy().astype(bool).reshape(nv, 12)
        verts_h = host[1].numpy().T
        keep = valid_h.any(axis=1)
        tris = triangulate(valid_h, ijk)
        flat = valid_h.ravel()
        data = DualContouringData(host[2].numpy().T[flat], valid_h, host[3].numpy().T[flat])
        meshes.append(DualContouringMesh(verts_h[keep], tris, data))
    return meshes


def _raw_arrays(sol: Solutions, levels_host, grid: EngineGrid, options, meshes) -> RawArraysSolution:
    """RawArraysSolution over the level-0 outputs (dense grid) or the octree -> regular fill; lazy."""
    ra = RawArraysSolution()
    first = sol.octrees_output[0]
    outs = first.outputs_centers
    last = outs[-1]
    g0 = first.grid_centers
    fb = lambda: last.combined_scalar_field.final_block
    fa = lambda: last.combined_scalar_field.faults_block
    lith = lambda h: h["lith"].get()
    faul = lambda h: h["faults"].get()
    sl = None
    if options.block_solutions_type == BlockSolutionType.DENSE_GRID and grid.dense_grid is not None:
        sl = g0.dense_grid_slice
        ra.set_lazy("lith_block", lambda: np.rint(fb()[sl]))
        ra.set_lazy("fault_block", lambda: np.rint(fa()[sl]))
    elif options.block_solutions_type == BlockSolutionType.OCTREE:
        base = grid.octree_grid.regular_grid_shape
        sl = slice(0, int(np.prod(base)))
        ra.set_lazy("lith_block", lambda: _fill_regular_from_octree(levels_host, base, lith))
        ra.set_lazy("fault_block", lambda: _fill_regular_from_octree(levels_host, base, faul))
    if sl is not None:
        ra.set_lazy("scalar_field_matrix", lambda: np.stack([o.exported_fields.scalar_field_everywhere[sl] for o in outs]))
        ra.set_lazy("block_matrix", lambda: np.stack([o.scalar_fields.values_block[0, sl] for o in outs]))
        ra.set_lazy("mask_matrix", lambda: np.stack([o.scalar_fields.mask_components[sl] for o in outs]))
        ra.set_lazy("mask_matrix_squeezed", lambda: np.stack([o.combined_scalar_field.squeezed_mask_array[sl] for o in outs]))

        def litho_faults():
            lb, fbk = ra.lith_block, ra.fault_block
            return lb + fbk * max(len(np.unique(lb)), 1)
        ra.set_lazy("litho_faults_block", litho_faults)
    for name, attr in (("custom", "custom_grid_slice"), ("topography", "topography_slice"), ("sections", "sections_slice")):
        s_ = getattr(g0, attr)
        if s_.stop > s_.start:
            ra.set_lazy(name, lambda s_=s_: np.rint(fb()[s_]))
    if meshes is not None:
        ra.vertices = [m.vertices for m in meshes]
        ra.edges = [m.edges for m in meshes]
    return ra
