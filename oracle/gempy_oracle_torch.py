"""PyTorch-CPU float64 variant of the oracle's field + gradient evaluation -- TEST / BASELINE INFRASTRUCTURE ONLY.

north_star asks for the reference's numpy AND PyTorch-CPU backends to be timed on the host cores (the reference switches
backends at /root/reference/gempy/API/compute_API.py:45-50 via BackendTensor.change_backend_gempy).  The engine package is
absent from the reference tree, so this is the same restatement as ``oracle.gempy_oracle.evaluate`` written with torch
tensor ops (what the engine's PYTORCH backend does with its kernel matrices), multi-threaded through torch's intra-op
pool.  Only ``bench.py`` (cpu_baseline / --impl reference --backend torch) and ``tests/`` import it."""
from __future__ import annotations

import numpy as np
import torch

from . import gempy_oracle as orc


def _kernel_terms(r, a, kind):
    kind = orc._kernel_name(kind)
    if kind == "cubic":
        t = r / a
        t2 = t * t
        C = 1 - 7 * t2 + 35 / 4 * t2 * t - 7 / 2 * t2 * t2 * t + 3 / 4 * t2 * t2 * t2 * t
        Cp_r = (-14 + 105 / 4 * t - 35 / 2 * t2 * t + 21 / 4 * t2 * t2 * t) / a ** 2
        Cpp = 7 * (9 * t2 * t2 * t - 20 * t2 * t + 15 * t - 4) / (2 * a ** 2)
        return C, Cp_r, Cpp
    if kind == "exponential":
        e = torch.exp(-(r * r) / (2 * a * a))
        return e, -e / a ** 2, e * (r * r / a ** 4 - 1 / a ** 2)
    if kind == "matern_5_2":
        s = np.sqrt(5.0) * r / a
        e = torch.exp(-s)
        return (1 + s + s * s / 3) * e, -(5.0 / (3 * a * a)) * (1 + s) * e, -(5.0 / (3 * a * a)) * (1 + s - s * s) * e
    raise ValueError(kind)


def _dist(x, p):
    h = x[:, None, :] - p[None, :, :]
    return h, torch.sqrt((h * h).sum(-1) + orc.DIST_EPS)


def evaluate(st, ko, w, xyz, gradient: bool = True, chunk_elems: int = 500_000):
    """Same contract as oracle.gempy_oracle.evaluate for fault-free stacks with universal degree 0 / 1."""
    a, c_o, gi, ires, kind = ko.range, ko.c_o, ko.gi_res, ko.i_res, ko.kernel_function
    n_o, n_r = st.n_o, st.n_rest
    nu = orc.n_drift_terms(ko.uni_degree)
    if nu not in (0, 3):
        raise NotImplementedError("torch baseline: universal degree 0 or 1")
    T = lambda v: torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float64)
    w = T(w)
    w_g = w[:3 * n_o].reshape(3, n_o)
    w_i = w[3 * n_o:3 * n_o + n_r]
    mu = w[3 * n_o + n_r:3 * n_o + n_r + nu]
    ori, rest, ref = T(st.ori_pos), T(st.rest), T(st.ref)
    xyz = T(np.asarray(xyz, float).reshape(-1, 3))
    m = xyz.shape[0]
    Z = torch.zeros(m, dtype=torch.float64)
    G = torch.zeros((m, 3), dtype=torch.float64) if gradient else None
    step = max(1, int(chunk_elems // max(1, (3 * n_o + 2 * n_r))))
    for s in range(0, m, step):
        x = xyz[s:s + step]
        z = torch.zeros(x.shape[0], dtype=torch.float64)
        g = torch.zeros((x.shape[0], 3), dtype=torch.float64) if gradient else None
        if n_o:
            h, r = _dist(x, ori)
            _, kp, ka = _kernel_terms(r, a, kind)
            hw = torch.einsum("moa,ao->mo", h, w_g)
            z += c_o * gi * (-(hw * kp)).sum(1)
            if gradient:
                Tt = (kp - ka) / (r * r + orc.REG_EPS)
                g += c_o * (torch.einsum("moa,mo->ma", h, hw * Tt) - torch.einsum("mo,ao->ma", kp, w_g))
        if n_r:
            h1, r1 = _dist(x, rest)
            h0, r0 = _dist(x, ref)
            C1, kp1, _ = _kernel_terms(r1, a, kind)
            C0, kp0, _ = _kernel_terms(r0, a, kind)
            z += c_o * ires * ((C1 - C0) @ w_i)
            if gradient:
                g += c_o * gi * (torch.einsum("mia,mi->ma", h1, kp1 * w_i) - torch.einsum("mia,mi->ma", h0, kp0 * w_i))
        if nu:
            z += gi * (x @ mu)
            if gradient:
                g += mu[None, :]
        Z[s:s + step] = z
        if gradient:
            G[s:s + step] = g
    return (Z.numpy(), G.numpy()) if gradient else Z.numpy()
