"""Forward gravity: the step downstream of the lithology block.

``gravity[c] = sum_k tz[k] * density[lith_id(centre_c + kernel_voxel_k)]`` on the centered grid GemPy's bridge builds
(gempy/modules/data_manipulation/_engine_factory.py:82-87).  The kernel geometry and the vertical gravity gradient
``tz`` of a voxel (Plouff's prism formula) restate what ``gp.calculate_gravity_gradient`` (exported at
gempy/API/__init__.py:65) computes; the restatement is pinned by the reference's known answer
``solutions.gravity == [-1624.1714]`` (4 decimals, test/test_modules/test_geophysics/test_gravity.py:67-89), which this
module reproduces (tests/test_oracle.py::test_gravity_known_answer, tests/test_gpu_parity.py::test_gravity_known_answer_gpu).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


def centered_grid_kernel(resolution, radius):
    """Kernel voxel centres and half-extents around a device at the origin: geometric spacing, both signs in x and y,
    downwards only in z (shifted by 5 % and stretched by 1.2)."""
    radius = np.broadcast_to(np.asarray(radius, dtype=np.float64), (3,))
    g2, d = [], []
    for ax in range(3):
        if ax == 2:
            g = np.geomspace(0.01, 1, int(resolution[ax]))
            g2.append((np.concatenate(([0.0], g)) + 0.05) * -radius[ax] * 1.2)
        else:
            g = np.geomspace(0.01, 1, int(resolution[ax] / 2))
            g2.append(np.concatenate((-g[::-1], [0.0], g)) * radius[ax])
        d.append(np.diff(np.pad(g2[ax], 1, "reflect", reflect_type="odd")))
    mesh = np.meshgrid(*g2)
    left = np.meshgrid(d[0][:-1] / 2, d[1][:-1] / 2, d[2][:-1] / 2)
    right = np.meshgrid(d[0][1:] / 2, d[1][1:] / 2, d[2][1:] / 2)
    flat = lambda m: np.vstack([a.ravel() for a in m]).T.astype(np.float64)
    return flat(mesh), flat(left), flat(right)


def calculate_gravity_gradient(centered_grid, ugal: bool = True) -> np.ndarray:
    """tz of every kernel voxel (same for every device centre)."""
    c = centered_grid.kernel_grid_centers
    dl, dr = centered_grid.kernel_dxyz_left, centered_grid.kernel_dxyz_right
    corners = [np.stack((c[:, a] - dl[:, a], c[:, a] + dr[:, a]), axis=1) for a in range(3)]
    x = np.repeat(corners[0], 4, axis=1)
    y = np.tile(np.repeat(corners[1], 2, axis=1), (1, 2))
    z = np.tile(corners[2], (1, 4))
    s = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    mu = np.array([1, -1, -1, 1, -1, 1, 1, -1])
    G = 6.674e-3 if ugal else 6.67428e-11
    return G * np.sum(-1 * mu * (x * np.log(y + s) + y * np.log(x + s) - z * np.arctan(x * y / (z * s))), axis=1)


@dataclass
class GravityInput:
    tz: np.ndarray
    densities: np.ndarray


@dataclass
class GeophysicsInput:
    """GeophysicsInput(tz, densities) as the reference builds it (test_gravity.py:76-79)."""
    tz: Optional[np.ndarray] = None
    densities: Optional[np.ndarray] = None

    def __post_init__(self):
        if self.tz is not None:
            self.tz = np.ascontiguousarray(self.tz, dtype=np.float64)
        if self.densities is not None:
            self.densities = np.ascontiguousarray(self.densities, dtype=np.float64)
