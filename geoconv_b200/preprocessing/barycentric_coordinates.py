"""Build-time helper of the hot path: the polar template grid.

Mirrors ``geoconv.preprocessing.barycentric_coordinates.create_template_matrix``
(/root/reference/src/geoconv/preprocessing/barycentric_coordinates.py:91-124).  The rest of that
reference module (GPC systems -> barycentric coordinates) is offline preprocessing and out of
scope.
"""
import numpy as np


def create_template_matrix(n_radial, n_angular, radius, in_cart=False):
    """Template vertices K[j, k] = (rho, theta) with rho = (j+1)*radius/n_radial and
    theta = 2*pi*(k+1)/n_angular (so the last angular coordinate is 2*pi == 0).

    Returns a float64 array (n_radial, n_angular, 2); with ``in_cart`` the entries are
    (x, y) = (rho*cos(theta), rho*sin(theta)) instead.
    """
    rho = np.arange(1, n_radial + 1, dtype=np.float64) * radius / n_radial
    theta = 2.0 * np.arange(1, n_angular + 1, dtype=np.float64) * np.pi / n_angular
    grid = np.empty((n_radial, n_angular, 2), dtype=np.float64)
    grid[..., 0] = rho[:, None]
    grid[..., 1] = theta[None, :]
    if in_cart:
        grid = np.stack([grid[..., 0] * np.cos(grid[..., 1]), grid[..., 0] * np.sin(grid[..., 1])], axis=-1)
    return grid
