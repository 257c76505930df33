"""Host side of the B200 backend: ``compute_model`` with the engine's signature.

Mirrors ``gempy_engine.compute_model(interpolation_input, options, data_descriptor, geophysics_input)``
(the call GemPy makes at /root/reference/gempy/API/compute_API.py:68-73) and returns a ``Solutions`` object
with the attributes GemPy reads (gempy/core/data/geo_model.py:100-127).  The pipeline per octree level is

    host:       level buffers (centres emitted by the previous level, unique corners through a lattice hash, surface points)
    one C call: per stack [level 0: fault values at the stack's points -> lower-triangle covariance -> symmetric solve
                (pivoted LU for tiny / indefinite systems) -> packed weights -> isovalues]
                -> ONE fused launch per segment: field (+gradient) + fault drift + activator (octet / point-list / z-run kernel)
                then corner segment <- unique corners, masks + combination -> lith / fault blocks       (gpb_model_run_level)
    next level: corner-id refinement test -> count (the level's one synchronisation) -> children
    finally:    dual contouring on the surface level (gpb_dual_contour, no synchronisation), lazy host containers

Every arithmetic step is a CUDA kernel of libgempy_b200.so called through the C ABI (gempy_b200/_lib.py).
torch is used for device-memory ownership, streams and (multi-GPU) torch.distributed only.
There is no CPU fallback: without the library or a CUDA device this module raises.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import warnings
import weakref
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .comm import Comm
from .data import (BlockSolutionType, CombinedScalarFieldsOutput, Deferred, DualContouringData, DualContouringMesh, EngineGrid,
                   ExportedFields, GenericGrid, InputDataDescriptor, InterpolationInput, InterpolationOptions,
                   InterpOutput, MeshExtractionMaskingOptions, OctreeLevel, RawArraysSolution, RegularGrid, ScalarFieldOutput,
                   Solutions, StackRelationType)

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
    """Per-stack device tables plus everything the level executor needs, allocated once per ``compute_model`` call:
    model-wide constants (all surface points, unit ids), per-stack outputs of the solve (weights, packed evaluation table,
    isovalues), the fault-drift tables the executor fills, and the native model handle (gpb_model_create).  No
    host-to-device copy is issued inside the per-level loop."""

    def __init__(self, eng: "B200Engine", ii: InterpolationInput, desc: InputDataDescriptor, options: InterpolationOptions):
        ko = options.kernel_options
        ss = desc.stack_structure
        n_st = ss.n_stacks
        device = eng.device
        super().__init__(StackTables(ii, desc, i, ko, device) for i in range(n_st))
        self.eng = eng
        self.sp_all = torch.as_tensor(np.ascontiguousarray(ii.surface_points.sp_coords.T), dtype=F64, device=device)
        self.unit_values = torch.as_tensor(np.asarray(ii.unit_values, dtype=np.float64), device=device)
        self.rel = [_rel_code(r) for r in ss.masking_descriptor]
        self.iso_min = eng.empty(n_st)
        self.iso_max = eng.empty(n_st)
        self.fault_min = eng.empty(n_st)
        self.iso_all = eng.empty(int(ss.number_of_surfaces_per_stack.sum()))
        self._iso_host = None
        self.weights: List[torch.Tensor] = []
        self.eval_tables: List[torch.Tensor] = []
        self.isovalues: List[torch.Tensor] = []
        arr = (_lib.GpbModelStack * n_st)()
        self._keep = []
        for i, st in enumerate(self):
            n_f = int(st.active_faults.size)
            st.n_faults = n_f
            if n_f:
                st.fault_rest = eng.empty(n_f, max(st.n_rest, 1))
                st.fault_ref = eng.empty(n_f, max(st.n_rest, 1))
                host_ids = (C.c_int * n_f)(*[int(g) for g in st.active_faults])
                dev_ids = torch.as_tensor(np.asarray(st.active_faults, dtype=np.int32), device=device)
                self._keep += [host_ids, dev_ids]
            st._struct = None
            sct = st.struct()
            ids = self.unit_values[st.surf_slice.start:st.surf_slice.stop + 1]
            if ids.shape[0] != st.n_surf + 1:
                raise ValueError("unit_values must hold one id per surface plus the basement")
            if st.n_surf > 64:
                raise ValueError(f"stack {i}: more than 64 surfaces in one stack")
            w = eng.empty(st.n)
            tab = eng.empty(int(eng.lib.gpb_eval_table_doubles(C.byref(sct))))
            iso = self.iso_all[st.surf_slice.start:st.surf_slice.stop]
            self.weights.append(w)
            self.eval_tables.append(tab)
            self.isovalues.append(iso)
            ms = arr[i]
            ms.st = sct
            ms.relation = self.rel[i]
            ms.fault_stacks_host = host_ids if n_f else None
            ms.fault_stacks_dev = _ptr(dev_ids) if n_f else None
            ms.sp_begin = st.sp_slice.start
            ms.n_sp = st.sp_slice.stop - st.sp_slice.start
            ms.unit_ids = _ptr(ids)
            ms.weights = _ptr(w)
            ms.eval_table = _ptr(tab)
            ms.isovalues = _ptr(iso)
        solver = 1 if os.environ.get("GPB_SOLVER", "sym") == "lu" else 0
        d = _lib.GpbModelDesc(n_st, arr, _ptr(self.sp_all), int(self.sp_all.shape[1]), float(options.sigmoid_slope),
                              _ptr(self.iso_min), _ptr(self.iso_max), _ptr(self.fault_min), solver)
        self.handle = C.c_void_p()
        _lib.check(eng.lib.gpb_model_create(C.byref(d), C.byref(self.handle)))
        self._destroy = eng.lib.gpb_model_destroy

    def iso_host(self, i: int) -> np.ndarray:
        """Isovalues of stack i on the host (one device-to-host copy for the whole model, on first use)."""
        if self._iso_host is None:
            self._iso_host = self.iso_all.cpu().numpy()
        return self._iso_host[self[i].surf_slice.start:self[i].surf_slice.stop]

    def solver_paths(self) -> List[str]:
        names = {0: "none", 1: "sym", 2: "lu"}
        return [names.get(int(self.eng.lib.gpb_model_solver_path(self.handle, i)), "?") for i in range(len(self))]

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            self._destroy(h)
            self.handle = None


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

    def seg_offset(self, name: str) -> int:
        off = 0
        for s in self.segments:
            if s.name == name:
                return off
            off += s.m
        return off

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
        """The caller's current CUDA stream on this engine's device.  Inside ``hold_stream()`` the lookup is done once."""
        if self._held_stream is not None:
            return self._held_stream
        return torch.cuda.current_stream(self.device).cuda_stream

    @contextlib.contextmanager
    def hold_stream(self):
        """Makes this engine's device the current CUDA device (the library launches on the current device) and caches the
        stream handle for the duration of a call."""
        outer = self._held_stream
        with torch.cuda.device(self.device):
            if outer is None:
                self._held_stream = torch.cuda.current_stream(self.device).cuda_stream
            try:
                yield
            finally:
                self._held_stream = outer

    def empty(self, *shape, dtype=F64) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # -- stages ----------------------------------------------------------------------------------------------
    def assemble(self, st: StackTables, extra_rows: int = 0, lower_only: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """System matrix (column-major: tensor row j = matrix column j) and right-hand side.  With ``extra_rows`` the
        leading dimension is n + extra_rows rounded up to an even number (16-byte aligned columns; the symmetric solve
        carries its right-hand sides as extra rows) and the tensor has shape (n, lda).  ``lower_only``: systems of order
        >= 512 get only their lower triangle written (all the symmetric solve reads)."""
        n = st.n
        lda = n if extra_rows == 0 else (n + extra_rows + 1) & ~1
        A = self.empty(n, lda)              # symmetric at this point
        b = self.empty(n)
        s = st.struct()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gpb_assemble_cov_ex(C.byref(s), _ptr(A), lda, _ptr(b), 1 if lower_only else 0, self.stream))
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
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gpb_lu_solve(n, _ptr(A), lda, _ptr(b), 1, n, _ptr(ipiv), _ptr(info), self.stream))
        self._check_info(info, what)
        return b

    SMALL_N = 160          # systems up to this order are solved by the one-CTA LU (every reference example model)

    def solve_stack(self, st: StackTables, what: str = "stack") -> Tuple[torch.Tensor, str]:
        """Assemble and solve one stack's saddle-point system; returns (weights, path).  Systems larger than SMALL_N
        take the symmetric path (Cholesky of the covariance block + Schur complement of the drift rows,
        gpb_sym_solve); if the covariance block is not numerically positive definite the system is re-assembled and
        solved by the pivoted LU.  (The level executor does the same natively, gpb_model_solve_stack.)"""
        n = st.n
        nk = 3 * st.n_ori + st.n_rest
        if n > self.SMALL_N and nk >= 1 and os.environ.get("GPB_SOLVER", "sym") != "lu":
            A, b = self.assemble(st, extra_rows=1, lower_only=True)
            info = torch.zeros(1, dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
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
        with torch.cuda.device(self.device):
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
        with torch.cuda.device(self.device):
            if seg.grid is not None:
                _lib.check(self.lib.gpb_eval_regular(C.byref(s), _ptr(src), C.byref(seg.grid), seg.i0, seg.i0 + seg.m,
                                                     fv, L, _ptr(Z, o8), gp[0], gp[1], gp[2], self.stream))
            else:
                _lib.check(self.lib.gpb_eval_points(C.byref(s), _ptr(src), _ptr(seg.xyz), seg.xyz.stride(0), seg.m,
                                                    fv, L, _ptr(Z, o8), gp[0], gp[1], gp[2], self.stream))

    # -- all stacks on one domain: the level executor ------------------------------------------------------------
    def run_level(self, tables: ModelTables, eval_segs: List[tuple], out_parts: List[Tuple[str, int]], solve: bool,
                  gradient: bool, comm: Optional[Comm] = None, hidden: int = 0,
                  expand: Optional[Tuple[torch.Tensor, int, int, int]] = None) -> FieldsOnDevice:
        """Evaluate every stack on one level.

        ``out_parts``: the level's output segments (name, count) in output order WITHOUT the surface-point tail, which
        always comes last; ``eval_segs``: what is evaluated, tuples ("points", count, out_offset, xyz_tensor_view) or
        ("regular", count, out_offset, (grid descriptor, first index)), optionally followed by a device int64 tensor with the
        actual point count (<= count) -- the surface-point tail must be covered by one of them.  ``hidden`` extra output columns beyond the official length hold the de-duplicated corners; ``expand`` =
        (map, src_offset, dst_offset, count) fills the corner segment from them before the combination.
        With a multi-rank ``comm`` the stacks are walked one by one and every fault block's minimum is all-reduced before
        the next stack needs it; otherwise one native call does the level."""
        comm = comm or Comm()
        n_st = len(tables)
        n_sp = int(tables.sp_all.shape[1])
        L = sum(c for _, c in out_parts) + n_sp
        ld = L + hidden
        Z = self.empty(n_st, ld)
        G = self.empty(n_st, 3, ld) if gradient else None
        block = self.empty(n_st, ld)
        final_block = self.empty(ld)
        faults_block = self.empty(ld)
        squeezed = self.empty(n_st, ld, dtype=torch.uint8)
        mask = self.empty(n_st, ld, dtype=torch.uint8)
        arr = (_lib.GpbSegment * len(eval_segs))()
        keep = []
        for k, seg in enumerate(eval_segs):
            kind, cnt, off, what = seg[:4]
            arr[k].count, arr[k].out_offset = int(cnt), int(off)
            if len(seg) > 4 and seg[4] is not None:           # device-side point count (compacted list)
                arr[k].count_dev = _ptr(seg[4])
                keep.append(seg[4])
            if kind == "regular":
                arr[k].kind, arr[k].grid, arr[k].i0 = _lib.GPB_SEG_REGULAR, what[0], int(what[1])
            else:
                arr[k].kind = _lib.GPB_SEG_OCTETS if kind == "octets" else _lib.GPB_SEG_POINTS
                arr[k].xyz, arr[k].ld_xyz = _ptr(what), int(what.stride(0))
                keep.append(what)
        lvl = _lib.GpbLevel(ld, len(eval_segs), arr, L - n_sp, _ptr(Z), _ptr(G), _ptr(block), _ptr(final_block), _ptr(faults_block),
                            _ptr(squeezed), _ptr(mask), None, 0, 0, 0, L)
        if expand is not None:
            lvl.expand_map, lvl.expand_src, lvl.expand_dst, lvl.expand_count = _ptr(expand[0]), int(expand[1]), int(expand[2]), int(expand[3])
            keep.append(expand[0])
        lib, h, stream = self.lib, tables.handle, self.stream
        ws, ws_bytes = None, 0
        if solve:                                         # scratch of the largest system, owned by torch's allocator
            ws_bytes = int(lib.gpb_model_workspace_bytes(h))
            ws = self.empty((ws_bytes + 7) // 8)
        if comm.world == 1:
            _lib.check(lib.gpb_model_run_level(h, C.byref(lvl), int(solve), _ptr(ws), ws_bytes, stream))
        else:
            for i in range(n_st):
                if solve:
                    _lib.check(lib.gpb_model_solve_stack(h, i, C.byref(lvl), None, _ptr(ws), ws_bytes, stream))
                _lib.check(lib.gpb_model_eval_stack(h, i, C.byref(lvl), stream))
                if tables.rel[i] == StackRelationType.FAULT.value:
                    comm.all_reduce_min(tables.fault_min[i:i + 1])
            _lib.check(lib.gpb_model_combine(h, C.byref(lvl), stream))
        out_segments = [Segment(nm, c) for nm, c in out_parts]
        f = FieldsOnDevice(out_segments, L - n_sp, Z[:, :L], None if G is None else G[:, :, :L], block[:, :L], final_block[:L],
                           faults_block[:L], squeezed[:, :L], mask[:, :L], tables.isovalues, tables.weights, [None] * n_st,
                           tables.eval_tables)
        f._keep = keep
        return f

    def gradient_at(self, st: StackTables, src: torch.Tensor, xyz: torch.Tensor) -> torch.Tensor:
        """Engine-convention gradient [3, m] of one stack's field at explicit points.  The fault drift has no
        gradient term, so the fault columns are skipped."""
        m = xyz.shape[1]
        Z = self.empty(m)
        G = self.empty(3, m)
        s = _lib.GpbStack.from_buffer_copy(st.struct())
        s.n_faults = 0
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gpb_eval_points(C.byref(s), _ptr(src), _ptr(xyz), xyz.stride(0), m, None, 0, _ptr(Z), _ptr(G[0]),
                                                _ptr(G[1]), _ptr(G[2]), self.stream))
        return G

    # -- octree ----------------------------------------------------------------------------------------------
    def corners_into(self, centers: torch.Tensor, nv: int, d: np.ndarray, out: torch.Tensor):
        """8 corners per voxel of centers[:, :nv] (row stride respected) written to out[:, :8 nv]."""
        if nv:
            _lib.check(self.lib.gpb_voxel_corners(_ptr(centers), centers.stride(0), nv, d[0] / 2, d[1] / 2, d[2] / 2, _ptr(out),
                                                  out.stride(0), self.stream))

    def corners_of(self, centers: torch.Tensor, d: np.ndarray) -> torch.Tensor:
        nv = centers.shape[1]
        out = self.empty(3, 8 * nv)
        with torch.cuda.device(self.device):
            self.corners_into(centers, nv, d, out)
        return out

    def mark(self, nv: int, lith_corners: torch.Tensor, fault_corners: torch.Tensor, force_all: bool) -> torch.Tensor:
        """Refinement test on nv voxels given the ids at their 8 corners each (uint8 marks)."""
        mark = self.empty(nv, dtype=torch.uint8)
        if nv:
            _lib.check(self.lib.gpb_mark_voxels(_ptr(lith_corners), _ptr(fault_corners), nv, int(force_all), _ptr(mark),
                                                self.stream))
        return mark

    def count_marked(self, mark: torch.Tensor) -> Tuple[int, torch.Tensor]:
        """Number of marked voxels (the one host synchronisation of a level: the next level must be sized) and the scan
        offsets the emission re-uses."""
        nv = mark.shape[0]
        offsets = self.empty(int(self.lib.gpb_scan_elems(nv)), dtype=torch.int64)
        n = C.c_longlong(0)
        _lib.check(self.lib.gpb_count_marked(_ptr(mark), nv, _ptr(offsets), C.byref(n), self.stream))
        return int(n.value), offsets

    def emit_into(self, centers: torch.Tensor, nv: int, d: np.ndarray, mark: torch.Tensor, offsets: torch.Tensor,
                  out: torch.Tensor):
        """Children (8 per marked voxel, parent order preserved) of centers[:, :nv] written to the first columns of out."""
        if nv:
            _lib.check(self.lib.gpb_emit_marked(_ptr(centers), centers.stride(0), nv, _ptr(mark), _ptr(offsets), d[0] / 4, d[1] / 4,
                                                d[2] / 4, _ptr(out), out.stride(0), self.stream))

    def copy_rows(self, dst: torch.Tensor, src: torch.Tensor):
        """dst[r, :] = src[r, :] for 2-D float64 device tensors with unit column stride (copy engine, no kernel)."""
        rows, cols = src.shape
        _lib.check(self.lib.gpb_copy_2d(_ptr(dst), dst.stride(0), _ptr(src), src.stride(0), rows, cols, self.stream))

    def gather_fields(self, f: FieldsOnDevice, totals: Sequence[int], comm: Comm) -> FieldsOnDevice:
        """All-gather a range-sharded level into whole arrays (segment by segment; the surface-point tail is
        replicated).  Identity on a single rank."""
        if comm.world == 1:
            return f
        n_sp = f.Z.shape[1] - f.grid_size

        L_tot = int(sum(totals)) + n_sp

        def full(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
            if t is None:
                return None
            out = torch.empty(*t.shape[:-1], L_tot, dtype=t.dtype, device=t.device)
            off = o_t = 0
            for seg, tot in zip(f.segments, totals):          # each segment's shards land in place, row by row
                comm.all_gather_rows_into(out[..., o_t:o_t + int(tot)], t[..., off:off + seg.m], int(tot))
                off += seg.m
                o_t += int(tot)
            out[..., o_t:].copy_(t[..., off:off + n_sp])
            return out

        segs = [Segment(sg.name, int(tot)) for sg, tot in zip(f.segments, totals)]
        return FieldsOnDevice(segs, int(sum(totals)), full(f.Z), full(f.G), full(f.block), full(f.final_block),
                              full(f.faults_block), full(f.squeezed), full(f.mask), f.isovalues, f.weights, f.cond, f.srcs)


# ------------------------------------------------------------------------------------------------ dense field API
def compute_dense_fields(interpolation_input: InterpolationInput, options: InterpolationOptions,
                         data_descriptor: InputDataDescriptor, *, stack: int = 0, engine: Optional[B200Engine] = None,
                         point_range: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None,
                         n_slabs: int = 16, device: Optional[int] = None, wave_aligned: bool = True) -> torch.Tensor:
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
    fr = desc.stack_structure.faults_relations
    if fr is not None and np.asarray(fr)[:, stack].any():
        raise ValueError("compute_dense_fields handles fault-free stacks; use compute_model for faulted ones")
    with eng.hold_stream():
        st = StackTables(ii, desc, stack, ko, eng.device)
        w, _ = eng.solve_stack(st, f"stack {stack}")
        src = eng.pack(st, w)
        if out is None:
            out = torch.empty((4, m), dtype=F64, pin_memory=True)
        gd = regular_descriptor(g)
        nyz = int(g.regular_grid_shape[1] * g.regular_grid_shape[2])
        # Slab size: a whole number of waves of the persistent evaluation kernel (one CTA per SM, 256 threads x 8 points per
        # chunk) -- a 4-plane slab of a 512^3 grid is 3.46 waves and would idle 13 % of every launch; smaller ranges fall
        # back to whole x planes.  Both are multiples of 8 points, which keeps every slab on the z-run kernel.
        per = max(1, -(-m // max(1, n_slabs)))
        wave = torch.cuda.get_device_properties(eng.device).multi_processor_count * 256 * 8
        if wave_aligned and per >= wave:
            per = -(-per // wave) * wave
        elif per > nyz:
            per = -(-per // nyz) * nyz
        compute = torch.cuda.current_stream(eng.device)
        if getattr(eng, "_copier", None) is None:
            eng._copier = torch.cuda.Stream(eng.device)          # one copy stream per engine, created once
        copier = eng._copier
        copier.wait_stream(compute)
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
_PINNED_MIN_BYTES = int(os.environ.get("GPB_PINNED_MIN_BYTES", str(4 << 20)))


def _np(t: Optional[torch.Tensor]) -> Optional[np.ndarray]:
    """Device tensor -> numpy.  Row-strided 2-D float64 views (slices of a level's buffers) are packed with a 2-D copy on
    the copy engine first, so that no elementwise kernel is launched for a read-back."""
    if t is None:
        return None
    t = t.detach()
    if t.is_cuda and not t.is_contiguous() and t.dim() == 2 and t.stride(1) == 1 and t.dtype == F64 and t.shape[1] > 0:
        tmp = torch.empty(t.shape, dtype=F64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().gpb_copy_2d(_ptr(tmp), tmp.stride(0), _ptr(t), t.stride(0), t.shape[0], t.shape[1],
                                              torch.cuda.current_stream(t.device).cuda_stream))
        t = tmp
    if t.is_cuda and t.numel() * t.element_size() >= _PINNED_MIN_BYTES and t.is_contiguous():
        # large read-backs go through page-locked memory (3-5x the bandwidth of a pageable copy); the numpy array is a
        # view of the pinned tensor, which torch's caching host allocator recycles once the array is dropped
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return host.numpy()
    return t.cpu().numpy()


def _np_rint(t: torch.Tensor) -> np.ndarray:
    """numpy.rint of a contiguous device vector, rounded on the device first (a host rint of a 512^3 block costs more than
    its read-back)."""
    if not t.is_cuda or t.numel() == 0:
        return np.rint(_np(t))
    t = t.detach().contiguous()
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().gpb_rint(_ptr(t), t.numel(), _ptr(out), torch.cuda.current_stream(t.device).cuda_stream))
    return _np(out)


def _level_outputs(f: FieldsOnDevice, grid: EngineGrid, rel_enum: Sequence, iso_host=None) -> List[InterpOutput]:
    """Host containers over the device results; every array is copied on first access only."""
    D = Deferred
    fb = D(lambda: _np(f.final_block))
    fa = D(lambda: _np(f.faults_block))
    outs = []
    for i in range(f.Z.shape[0]):
        g = (None, None, None) if f.G is None else tuple(D(lambda i=i, a=a: _np(f.G[i, a])) for a in range(3))
        iso = D(lambda i=i: iso_host(i).copy()) if iso_host is not None else D(lambda i=i: _np(f.isovalues[i]))
        ef = ExportedFields(D(lambda i=i: _np(f.Z[i])), g[0], g[1], g[2], f.grid_size, iso)
        sfo = ScalarFieldOutput(D(lambda i=i: _np(f.weights[i])), grid, ef, D(lambda i=i: _np(f.block[i])[None, :]),
                                rel_enum[i], D(lambda i=i: _np(f.mask[i]).astype(bool)))
        comb = CombinedScalarFieldsOutput(D(lambda i=i: _np(f.squeezed[i]).astype(bool)), fb, fa)
        outs.append(InterpOutput(sfo, comb))
    return outs


def _lattice(root: RegularGrid, level: int) -> _lib.GpbRegularGrid:
    """Voxel lattice of octree level `level` (0 = root): centre of cell (0, 0, 0) with the grid shift, cell size, cells."""
    e, s = root.orthogonal_extent, root.regular_grid_shape * (2 ** level)
    d = np.array([(e[1] - e[0]) / s[0], (e[3] - e[2]) / s[1], (e[5] - e[4]) / s[2]])
    return _lib.GpbRegularGrid(e[0] + d[0] / 2 + GRID_SHIFT, e[2] + d[1] / 2 + GRID_SHIFT, e[4] + d[2] / 2 + GRID_SHIFT,
                               d[0], d[1], d[2], int(s[0]), int(s[1]), int(s[2]))


def _fill_regular_from_octree(eng: B200Engine, levels_dev, root: RegularGrid, which: str) -> np.ndarray:
    """Dense array at the finest octree resolution, on the device: every level's lattice is the level above repeated
    twice per axis with the level's own voxels written over it (the engine's octree -> regular fill used by
    RawArraysSolution).  levels_dev: [(centers [3, nv] view, nv, FieldsOnDevice)]."""
    with eng.hold_stream():
        dense = None
        for lvl, (centers, nv, f) in enumerate(levels_dev):
            lat = _lattice(root, lvl)
            cur = eng.empty(lat.nx * lat.ny * lat.nz)
            if dense is not None:
                _lib.check(eng.lib.gpb_upsample2(_ptr(dense), lat.nx // 2, lat.ny // 2, lat.nz // 2, _ptr(cur), eng.stream))
            vals = f.final_block if which == "lith" else f.faults_block
            _lib.check(eng.lib.gpb_scatter_lattice(_ptr(centers), centers.stride(0), nv, C.byref(lat), _ptr(vals), 1, _ptr(cur),
                                                   eng.stream))
            dense = cur
        return _np(dense)


_UNSUPPORTED_WARNED = set()


def _warn_unsupported(options: InterpolationOptions, ii: InterpolationInput) -> None:
    """Engine options this backend accepts but does not act on: say so once instead of silently ignoring them."""
    eo = options.evaluation_options
    notes = []
    if float(getattr(eo, "octree_curvature_threshold", -1.0)) != -1.0:
        notes.append("evaluation_options.octree_curvature_threshold (refinement is by corner ids only)")
    if float(getattr(eo, "octree_error_threshold", 1.0)) != 1.0:
        notes.append("evaluation_options.octree_error_threshold (refinement is by corner ids only)")
    mo = getattr(eo, "mesh_extraction_masking_options", MeshExtractionMaskingOptions.INTERSECT)
    if getattr(mo, "name", mo) not in ("INTERSECT", 3, "RAW", 4, "DISJOINT", 2):
        notes.append("evaluation_options.mesh_extraction_masking_options = NOTHING (meshes are masked by the squeezed stack mask = INTERSECT)")
    if int(getattr(eo, "evaluation_chunk_size", 500_000)) != 500_000:
        notes.append("evaluation_options.evaluation_chunk_size (the kernel matrix is never materialised: nothing to chunk)")
    if ii.weights:
        notes.append("interpolation_input.weights (the direct solver needs no warm start: every call assembles and solves again, "
                     "so edited inputs can never meet stale weights; cache_mode has nothing to select)")
    for n in notes:
        if n not in _UNSUPPORTED_WARNED:
            _UNSUPPORTED_WARNED.add(n)
            warnings.warn("gempy_b200: option without effect in this backend: " + n, stacklevel=3)


# ------------------------------------------------------------------------------------------------ entry point
def compute_model(interpolation_input: InterpolationInput, options: InterpolationOptions,
                  data_descriptor: InputDataDescriptor, geophysics_input=None, *, device: Optional[int] = None,
                  engine: Optional[B200Engine] = None, comm: Optional[Comm] = None) -> Solutions:
    """Drop-in for ``gempy_engine.compute_model`` (same positional/keyword signature; the keyword-only extras
    select the CUDA device and, for one-process-per-GPU runs, the torch.distributed group).  Every rank returns
    the same, complete ``Solutions``.  Raises ``NotImplementedError`` for magnetics input and ``GpbError`` for a
    singular system."""
    comm = comm or Comm()
    if geophysics_input is not None and interpolation_input.grid.geophysics_grid is None:
        raise ValueError("geophysics_input needs a centered (geophysics) grid")
    if geophysics_input is not None and getattr(geophysics_input, "magnetics_input", None) is not None:
        raise NotImplementedError("magnetics is outside the B200 backend's scope")
    eng = engine or B200Engine(device)
    with eng.hold_stream():
        return _compute_model(eng, interpolation_input, options, data_descriptor, geophysics_input, comm)


def compute_model_at(interpolation_input: InterpolationInput, options: InterpolationOptions,
                     data_descriptor: InputDataDescriptor, at: np.ndarray, **kwargs) -> np.ndarray:
    """Engine-level mirror of ``gp.compute_model_at`` (gempy/API/compute_API.py:89-114): the custom grid becomes the only
    active extra grid (``set_custom_grid(..., reset=True)``, grid_API.py:91-96; the octree grid always rides along,
    _engine_factory.py:88-96), the model is computed, and the lithology ids at ``at`` (TRANSFORMED coordinates, like
    everything at the engine boundary) come back as ``raw_arrays.custom``.  Like the reference, this replaces
    ``interpolation_input.grid`` (side effect)."""
    g = interpolation_input.grid
    interpolation_input.grid = EngineGrid(octree_grid=g.octree_grid, custom_grid=GenericGrid(np.asarray(at, dtype=np.float64)))
    sol = compute_model(interpolation_input, options, data_descriptor, **kwargs)
    return sol.raw_arrays.custom


def _shard_pays(n_pts: int, n_src_total: int, n_stacks: int, world: int) -> bool:
    """Is a level worth sharding over `world` ranks?  Evaluation time saved (1e-12 s per point-source pair: 17 FP64
    instructions at the measured issue rate) against twice the cost of the exchanges it brings (the level's outputs --
    about 20 bytes per point and stack -- all-gathered at ~100 GB/s effective, plus 2 ms of per-stack launch sequences and
    small collectives).  Config 3 (5000 sources, one stack) and config 5 (25 000) shard; the 15-stack multi-fault model
    (125 sources per stack) does not -- sharded it ran 38 ms on 2 GPUs and 47 ms on 8 against 31 ms on one.
    GPB_SHARD_MIN_PAIRS=<pairs> replaces the model by a plain threshold (tests: 0 shards everything)."""
    env = os.environ.get("GPB_SHARD_MIN_PAIRS")
    if env is not None:
        return float(n_pts) * n_src_total >= float(env)
    t_eval = float(n_pts) * n_src_total * 1.0e-12
    t_xchg = float(n_pts) * n_stacks * 20.0 / 1.0e11 + 2.0e-3
    return t_eval * (1.0 - 1.0 / world) > 2.0 * t_xchg


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
    _warn_unsupported(options, ii)

    tables = ModelTables(eng, ii, desc, options)
    n_sp = int(tables.sp_all.shape[1])
    gradient = bool(eo.compute_scalar_gradient)
    n_levels = int(eo.number_octree_levels)
    dc_level = min(int(eo.number_octree_levels_surface), n_levels) - 1 if eo.mesh_extraction else -1
    root = grid.octree_grid
    d = root.dxdydz.copy()

    # explicit point sets of level 0 besides the octree centres (host tables; uploaded straight into the level's buffer)
    extras_host: List[Tuple[str, np.ndarray]] = []
    for name in ("custom_grid", "topography", "sections", "geophysics_grid"):
        g = getattr(grid, name)
        if g is not None:
            extras_host.append((name, np.ascontiguousarray(np.asarray(g.values, dtype=np.float64).T)))
    dense_desc = regular_descriptor(grid.dense_grid) if grid.dense_grid is not None else None

    octree_levels: List[OctreeLevel] = []
    levels_dev = []
    prev_regular = root
    dc_payload = None
    gravity = None
    dedupe = os.environ.get("GPB_NO_CORNER_DEDUPE") is None
    pending = None                                      # (parent centres, nv_parent, d_parent, mark, offsets) of the next emission
    nv = int(np.prod(root.regular_grid_shape))
    world_comm = comm
    n_src_total = sum(int(st.n_ori + st.n_rest + st.n_surf) for st in tables)
    for lvl in range(n_levels):
        need_corners = (lvl < n_levels - 1) or (lvl == dc_level)
        # a level is sharded over the ranks only when it is worth the exchanges (fault-minimum all-reduces, mark and output
        # all-gathers, one launch sequence per stack instead of one native call): otherwise every rank evaluates it whole
        comm = world_comm
        if world_comm.world > 1:
            n_pts = nv * (9 if need_corners else 1) + n_sp
            if lvl == 0:
                n_pts += (grid.dense_grid.n_points if grid.dense_grid is not None else 0) + sum(v.shape[1] for _, v in extras_host)
            if not _shard_pays(n_pts, n_src_total, n_st, world_comm.world):
                comm = Comm.solo()
        v0, v1 = comm.shard(nv)
        nvl = v1 - v0
        # ---- this level's voxel centres: [3, nv], the whole list on every rank (children are emitted by every rank)
        if lvl == 0:
            ax = root.axis_coords()
            gx, gy, gz = np.meshgrid(*ax, indexing="ij")
            centers_full = torch.as_tensor(np.stack([gx.ravel(), gy.ravel(), gz.ravel()]) + GRID_SHIFT, dtype=F64, device=eng.device)
        else:
            p_cen, p_nv, p_d, p_mark, p_off = pending
            centers_full = eng.empty(3, nv)
            eng.emit_into(p_cen, p_nv, p_d, p_mark, p_off, centers_full)
        centers_loc = centers_full[:, v0:v1]
        out_parts: List[Tuple[str, int]] = [("octree_grid", nvl)]
        totals = [nv]
        # below the root a level is a list of sibling octets (8 children per refined voxel, in emission order)
        octets = lvl > 0 and v0 % 8 == 0 and nvl % 8 == 0
        eval_segs: List[tuple] = [("octets" if octets else "points", nvl, 0, centers_loc)]
        off = nvl
        ex_loc = []
        if lvl == 0:
            if dense_desc is not None:
                i0, i1 = comm.shard(grid.dense_grid.n_points)
                out_parts.append(("dense_grid", i1 - i0))
                totals.append(grid.dense_grid.n_points)
                eval_segs.append(("regular", i1 - i0, off, (dense_desc, i0)))
                off += i1 - i0
            for name, vals in extras_host:
                i0, i1 = comm.shard(vals.shape[1])
                ex_loc.append(vals[:, i0:i1])
                out_parts.append((name, i1 - i0))
                totals.append(vals.shape[1])
        n_ex = sum(v.shape[1] for v in ex_loc)
        # ---- corners: de-duplicated (unique lattice corners evaluated once, the corner segment filled by a gather) or explicit
        n_u, cmap, scratch = 0, None, None
        use_dedupe = need_corners and dedupe and nvl > 0
        if use_dedupe:
            lat = _lattice(root, lvl)
            sb = int(eng.lib.gpb_corner_scratch_bytes(nvl))
            scratch = eng.empty((sb + 7) // 8)
            # no synchronisation: the count stays on the device; below the root a level is made of sibling octets with at most
            # 27 distinct corners per 8 voxels (+ slack for octets cut by a rank boundary), which bounds the buffers
            _lib.check(eng.lib.gpb_corner_unique_count(_ptr(centers_loc), centers_loc.stride(0), nvl, C.byref(lat), _ptr(scratch), sb,
                                                       None, eng.stream))
            n_u = 8 * nvl if lvl == 0 else min(8 * nvl, 27 * ((nvl + 7) // 8 + 2))
        n_cor_explicit = 8 * nvl if (need_corners and not use_dedupe) else 0
        pts = eng.empty(3, n_ex + n_cor_explicit + n_sp + n_u)          # explicit grids | (corners) | surface points | unique corners
        o = 0
        for vals in ex_loc:
            k = vals.shape[1]
            if k:
                eng.copy_rows(pts[:, o:o + k], torch.as_tensor(np.ascontiguousarray(vals), dtype=F64, device=eng.device))
            o += k
        if n_ex:
            eval_segs.append(("points", n_ex, off, pts[:, :n_ex]))
        off += n_ex
        if need_corners:
            out_parts.append(("corners", 8 * nvl))
            totals.append(8 * nv)
        corner_off = off
        if n_cor_explicit:
            eng.corners_into(centers_loc, nvl, d, pts[:, n_ex:n_ex + n_cor_explicit])
            # explicit grids and corners are contiguous in `pts` and in the outputs: one segment
            if n_ex:
                eval_segs.pop()
                off -= n_ex
            eval_segs.append(("points", n_ex + n_cor_explicit, off, pts[:, :n_ex + n_cor_explicit]))
            off += n_ex + n_cor_explicit
        elif need_corners:
            off += 8 * nvl
        sp_off = off                                                    # = L - n_sp
        eng.copy_rows(pts[:, n_ex + n_cor_explicit:n_ex + n_cor_explicit + n_sp], tables.sp_all)
        expand = None
        if use_dedupe:
            cmap = eng.empty(8 * nvl, dtype=torch.int32)
            tail = pts[:, n_ex + n_sp:]
            cnt_dev = eng.empty(1, dtype=torch.int64)                       # n_sp + number of unique corners, on the device
            _lib.check(eng.lib.gpb_corner_unique_emit(_ptr(centers_loc), centers_loc.stride(0), nvl, d[0] / 2, d[1] / 2, d[2] / 2,
                                                      _ptr(scratch), sb, _ptr(tail), tail.stride(0), _ptr(cmap), n_sp, _ptr(cnt_dev),
                                                      eng.stream))
            expand = (cmap, sp_off + n_sp, corner_off, 8 * nvl)
        # surface points and (behind them, beyond the official length) the unique corners: one segment
        eval_segs.append(("points", n_sp + n_u, sp_off, pts[:, n_ex + n_cor_explicit:], cnt_dev if use_dedupe else None))
        # ---- all stacks (one native call on a single rank)
        f_loc = eng.run_level(tables, eval_segs, out_parts, solve=(lvl == 0), gradient=gradient, comm=comm, hidden=n_u, expand=expand)
        f_loc._xyz = (centers_full, pts)                 # keeps the coordinate buffers alive with the fields
        if lvl == 0 and getattr(ko, "compute_condition_number", False):
            conds = []
            for i, st in enumerate(tables):
                A, _ = eng.assemble(st)
                conds.append(float(torch.linalg.cond(A).item()))
                del A
            f_loc.cond = conds
            ko.condition_number = conds[-1]
        # ---- refinement marks: local test, all-gathered so that every rank emits the identical child list
        mark_full = None
        if lvl < n_levels - 1:
            mark_loc = eng.mark(nvl, f_loc.final_block[corner_off:corner_off + 8 * nvl],
                                f_loc.faults_block[corner_off:corner_off + 8 * nvl], force_all=lvl < int(eo.octree_min_level))
            mark_full = comm.all_gather_cat(mark_loc, nv)
        f = eng.gather_fields(f_loc, totals, comm)
        # ---- host containers of this level
        centers_host = Deferred(lambda c=centers_full: _np(c).T.copy())     # explicit centres (shift included), as evaluated
        if lvl == 0:
            og0 = RegularGrid(root.orthogonal_extent, root.regular_grid_shape)
            og0._values, og0._n_points = centers_host, nv
            lvl_grid = EngineGrid(octree_grid=og0, dense_grid=grid.dense_grid, topography=grid.topography,
                                  sections=grid.sections, custom_grid=grid.custom_grid,
                                  geophysics_grid=grid.geophysics_grid)
        else:
            og = RegularGrid.from_octree_level(centers_host, prev_regular)
            og._dxdydz, og._n_points = d.copy(), nv
            prev_regular = og
            lvl_grid = EngineGrid(octree_grid=og)
        outs = _level_outputs(f, lvl_grid, rel_enum, tables.iso_host)
        level = OctreeLevel(grid_centers=lvl_grid, outputs_centers=outs,
                            _grid_corners=None if not need_corners else
                            Deferred(lambda c=centers_full, dd=d.copy(): EngineGrid.from_xyz_coords(_np(eng.corners_of(c, dd)).T)))
        level._device_fields = f
        octree_levels.append(level)
        levels_dev.append((centers_full, nv, f))
        if lvl == 0 and geophysics_input is not None:
            gravity = _forward_gravity(eng, geophysics_input, grid.geophysics_grid, f)
        if lvl == dc_level:
            dc_payload = (lvl, centers_full, d.copy(), eng.corners_of(centers_full, d), f)
        if lvl == n_levels - 1:
            break
        level._marked_voxels = Deferred(lambda mk=mark_full: _np(mk).astype(bool))
        n_marked, offsets = eng.count_marked(mark_full)
        pending = (centers_full, nv, d.copy(), mark_full, offsets)
        nv = 8 * n_marked
        d = d / 2

    meshes = None
    if eo.mesh_extraction and dc_payload is not None:
        meshes = _dual_contouring(eng, tables, dc_payload, root, getattr(eo, "mesh_extraction_masking_options", None))

    sol = Solutions(octree_levels, meshes, gravity, options.block_solutions_type)
    sol.raw_arrays = _raw_arrays(eng, sol, levels_dev, grid, options, meshes)
    sol._tables = tables            # device tables + the native model handle live as long as the lazy outputs
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
    block = f.final_block[sl]
    out = eng.empty(n_centers)
    _lib.check(eng.lib.gpb_gravity(_ptr(block), _ptr(dens_d), int(dens_d.shape[0]), _ptr(tz_d), n_centers, n_k, _ptr(out),
                                   eng.stream))
    return _np(out)


class _DeviceMesh:
    """Device buffers of one dual-contoured surface (worst-case sized) and its three counts; trimmed and copied to the
    host when the mesh is first read."""

    def __init__(self, nv, counts, verts, tris, valid, xyz_c, grad_c):
        self.nv, self.counts, self.verts, self.tris, self.valid, self.xyz_c, self.grad_c = nv, counts, verts, tris, valid, xyz_c, grad_c
        self._n = None

    def n(self):
        if self._n is None:
            self._n = [int(v) for v in self.counts.cpu().numpy()]
        return self._n

    def vertices(self) -> np.ndarray:
        V = self.n()[1]
        return _np(self.verts[:, :V]).T.copy()

    def triangles(self) -> np.ndarray:
        T = self.n()[2]
        return _np(self.tris[:T]).astype(np.int64)

    def dc_data(self) -> DualContouringData:
        E = self.n()[0]
        valid = _np(self.valid).astype(bool).reshape(self.nv, 12)
        return DualContouringData(_np(self.xyz_c[:, :E]).T.copy(), valid, _np(self.grad_c[:, :E]).T.copy())


def _dual_contouring(eng: B200Engine, tables: ModelTables, payload, root: RegularGrid, masking=None) -> List[DualContouringMesh]:
    """Every surface of every stack on the surface level, queued without a host round trip (gpb_dual_contour); the meshes
    are read back lazily.  ``masking`` (evaluation_options.mesh_extraction_masking_options): INTERSECT (default) keeps the
    voxels the stack owns at any of their corners (squeezed mask; fault stacks are never masked), RAW masks nothing,
    DISJOINT is not implemented (nor is it upstream)."""
    lvl, centers, d, corners, f = payload
    mname = getattr(masking, "name", masking)
    if mname in ("DISJOINT", 2):
        raise NotImplementedError("mesh_extraction_masking_options = DISJOINT is not implemented")
    raw = mname in ("RAW", 4)
    nv = int(centers.shape[1])
    c_off = f.seg_offset("corners")
    lat = _lattice(root, lvl)
    lib = eng.lib
    n_bytes = int(lib.gpb_dc_scratch_bytes(nv))
    scratch = eng.empty((n_bytes + 7) // 8)              # shared by all surfaces: the calls are ordered on one stream
    meshes: List[DualContouringMesh] = []
    for i, st in enumerate(tables):
        sct = st.struct()
        Zc = _ptr(f.Z[i], 8 * c_off)
        sq = None if (raw or tables.rel[i] == StackRelationType.FAULT.value) else _ptr(f.squeezed[i], c_off)
        for s_idx in range(st.n_surf):
            counts = eng.empty(3, dtype=torch.int64)
            verts = eng.empty(3, max(nv, 1))
            tris = eng.empty(max(6 * nv, 1), 3, dtype=torch.int32)
            valid = eng.empty(max(12 * nv, 1), dtype=torch.uint8)
            xyz_c = eng.empty(3, max(12 * nv, 1))
            grad_c = eng.empty(3, max(12 * nv, 1))
            _lib.check(lib.gpb_dual_contour(C.byref(sct), _ptr(tables.eval_tables[i]), _ptr(corners), corners.stride(0), Zc, sq,
                                            _ptr(centers), centers.stride(0), nv, _ptr(tables.isovalues[i], 8 * s_idx),
                                            C.byref(lat), 1.0, _ptr(scratch), n_bytes, _ptr(valid), _ptr(xyz_c), _ptr(grad_c),
                                            _ptr(verts), _ptr(tris), _ptr(counts), eng.stream))
            dm = _DeviceMesh(nv, counts, verts, tris, valid, xyz_c, grad_c)
            meshes.append(DualContouringMesh(Deferred(dm.vertices), Deferred(dm.triangles), Deferred(dm.dc_data)))
    return meshes


def _raw_arrays(eng: B200Engine, sol: Solutions, levels_dev, grid: EngineGrid, options, meshes) -> RawArraysSolution:
    """RawArraysSolution over the level-0 outputs (dense grid) or the octree -> regular fill; lazy."""
    ra = RawArraysSolution()
    first = sol.octrees_output[0]
    outs = first.outputs_centers
    last = outs[-1]
    g0 = first.grid_centers
    fb = lambda: last.combined_scalar_field.final_block
    fa = lambda: last.combined_scalar_field.faults_block
    sl = None
    if options.block_solutions_type == BlockSolutionType.DENSE_GRID and grid.dense_grid is not None:
        sl = g0.dense_grid_slice
        f0 = levels_dev[0][2]                         # rounded on the device, only the dense-grid slice is read back
        ra.set_lazy("lith_block", lambda: _np_rint(f0.final_block[sl]))
        ra.set_lazy("fault_block", lambda: _np_rint(f0.faults_block[sl]))
    elif options.block_solutions_type == BlockSolutionType.OCTREE:
        root = grid.octree_grid
        sl = slice(0, int(np.prod(root.regular_grid_shape)))
        ra.set_lazy("lith_block", lambda: _fill_regular_from_octree(eng, levels_dev, root, "lith"))
        ra.set_lazy("fault_block", lambda: _fill_regular_from_octree(eng, levels_dev, root, "faults"))
    if sl is not None:
        # [n_stacks, points] matrices: the slice is packed on the device and read back once (no host-side stack / copy)
        fd = levels_dev[0][2]
        ra.set_lazy("scalar_field_matrix", lambda: _np(fd.Z[:, sl]))
        ra.set_lazy("block_matrix", lambda: _np(fd.block[:, sl]))
        ra.set_lazy("mask_matrix", lambda: _np(fd.mask[:, sl]).astype(bool))
        ra.set_lazy("mask_matrix_squeezed", lambda: _np(fd.squeezed[:, sl]).astype(bool))

        ra_ref = weakref.ref(ra)          # no strong self-reference: a cycle would keep the level's device buffers alive
                                          # until the cyclic garbage collector runs (4.6 GB per octree-8 solution)

        def litho_faults():
            r = ra_ref()
            lb, fbk = r.lith_block, r.fault_block
            return lb + fbk * max(len(np.unique(lb)), 1)
        ra.set_lazy("litho_faults_block", litho_faults)
    for name, attr in (("custom", "custom_grid_slice"), ("topography", "topography_slice"), ("sections", "sections_slice")):
        s_ = getattr(g0, attr)
        if s_.stop > s_.start:
            ra.set_lazy(name, lambda s_=s_: np.rint(fb()[s_]))
    if meshes is not None:
        ra.set_lazy("vertices", lambda: [m.vertices for m in meshes])
        ra.set_lazy("edges", lambda: [m.edges for m in meshes])
    return ra
