"""Condition number of a stack's co-kriging system and the nugget optimiser built on it.

Mirror of ``gempy/modules/optimize_nuggets`` (``_optimizer.py:9-68``, ``_ops.py:6-109``; entry
``gempy/API/compute_API.py:117-134``): per structural group, an Adam loop (lr 0.01) lowers the condition number of the
group's covariance matrix by adjusting the surface-point nuggets; after every backward pass only the largest 1 % of the
gradient entries are kept (``_gradient_masking``, focus 0.01), nuggets are clamped at 1e-7, and the loop stops below the
target condition number or when the relative improvement drops under 1 % after ``patience`` epochs (``_has_converged``).

The reference differentiates ``torch.linalg.cond`` through the engine with autograd.  Here the matrix is assembled by
``gpb_assemble_cov`` and the derivative is analytic: the nuggets only enter the diagonal of the increment rows
(``diag_j += c_o (nugget_rest(j) + nugget_ref(j)) / 2``), the matrix is symmetric, so with the eigenpairs of largest and
smallest magnitude (lambda_M, q_M), (lambda_m, q_m)

    d cond / d diag_j = ( sign(lambda_M) q_M[j]^2 |lambda_m| - |lambda_M| sign(lambda_m) q_m[j]^2 ) / lambda_m^2 .

The symmetric eigen-decomposition is cuSOLVER's (``torch.linalg.eigh``): an auxiliary path, not the hot one.
Parity unpinned: the engine's own definition (which norm, which matrix) is not visible in the reference tree; the 2-norm
condition number of the full saddle-point matrix is used, as in ``oracle.gempy_oracle.condition_number_and_gradient``."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .compute import B200Engine, StackTables


def condition_number_and_gradient(eng: B200Engine, st: StackTables) -> Tuple[float, np.ndarray]:
    """-> (2-norm condition number of the stack's matrix, d cond / d nugget for the stack's surface points in input order)."""
    A, _ = eng.assemble(st)
    lam, Q = torch.linalg.eigh(A)
    mag = lam.abs()
    iM, im = int(torch.argmax(mag)), int(torch.argmin(mag))
    lM, lm = lam[iM], lam[im]
    cond = float(mag[iM] / mag[im])
    r0 = 3 * st.n_ori
    qM2 = Q[r0:r0 + st.n_rest, iM] ** 2
    qm2 = Q[r0:r0 + st.n_rest, im] ** 2
    g_diag = (torch.sign(lM) * qM2 * mag[im] - mag[iM] * torch.sign(lm) * qm2) / (lm * lm)
    half = 0.5 * float(st.ko.c_o) * g_diag
    n_sp = st.sp_slice.stop - st.sp_slice.start
    grad = torch.zeros(n_sp, dtype=torch.float64, device=A.device)
    grad.index_add_(0, st.rest_idx, half)
    grad.index_add_(0, st.ref_idx, half)
    return cond, grad.cpu().numpy()


def gradient_masking(grad: np.ndarray, focus: float = 0.01) -> np.ndarray:
    """Keep the ``int(n * focus)`` entries of largest magnitude, zero the rest (_ops.py:68-77)."""
    k = int(grad.size * focus)
    out = np.zeros_like(grad)
    if k > 0:
        top = np.argsort(-np.abs(grad), kind="stable")[:k]
        out[top] = grad[top]
    return out


def has_converged(current: float, previous: float, target: float = 1e5, epoch: int = 0, min_improvement: float = 0.01,
                  patience: int = 10) -> bool:
    """_ops.py:80-94."""
    if current < target:
        return True
    if epoch > patience:
        return abs(current - previous) / max(previous, 1e-8) < min_improvement
    return False


def optimize_nuggets(interpolation_input, options, data_descriptor, *, max_epochs: int = 10,
                     convergence_criteria: float = 1e5, only_stacks: Optional[Sequence[int]] = None, lr: float = 0.01,
                     patience: int = 10, min_impr: float = 0.01, focus: float = 0.01,
                     engine: Optional[B200Engine] = None, cond_and_grad=None) -> List[List[float]]:
    """Updates ``interpolation_input.surface_points.nugget_effect_scalar`` in place, stack by stack, and leaves the last
    condition number in ``options.kernel_options.condition_number``.  Returns the condition-number history per stack.
    ``cond_and_grad(stack_index) -> (cond, grad)`` replaces the device evaluation (tests drive the loop with the oracle)."""
    eng = None
    if cond_and_grad is None:
        eng = engine or B200Engine()
    ss = data_descriptor.stack_structure
    stacks = range(ss.n_stacks) if only_stacks is None else only_stacks
    ko = options.kernel_options
    history: List[List[float]] = []
    for i in stacks:
        sp0 = int(ss.number_of_points_per_stack[:i].sum())
        sp1 = sp0 + int(ss.number_of_points_per_stack[i])
        nug = torch.tensor(np.asarray(interpolation_input.surface_points.nugget_effect_scalar[sp0:sp1], dtype=np.float64))
        nug.requires_grad_(True)
        opt = torch.optim.Adam(params=[nug], lr=lr)
        prev, hist = float("inf"), []
        ko.optimizing_condition_number = True
        for epoch in range(max_epochs):
            opt.zero_grad()
            interpolation_input.surface_points.nugget_effect_scalar[sp0:sp1] = nug.detach().numpy()
            if cond_and_grad is not None:
                cur, g = cond_and_grad(i)
            else:
                cur, g = condition_number_and_gradient(eng, StackTables(interpolation_input, data_descriptor, i, ko, eng.device))
            nug.grad = torch.as_tensor(gradient_masking(np.asarray(g, dtype=np.float64), focus))
            opt.step()
            with torch.no_grad():
                nug.clamp_(min=1e-7)
            ko.condition_number = cur
            hist.append(cur)
            if has_converged(cur, prev, convergence_criteria, epoch, min_impr, patience):
                break
            prev = cur
        ko.optimizing_condition_number = False
        interpolation_input.surface_points.nugget_effect_scalar[sp0:sp1] = nug.detach().numpy()
        history.append(hist)
    return history
