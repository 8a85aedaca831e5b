"""Drop-in for ``geoconv.pytorch.layers.conv_chi_squared``
(/root/reference/src/geoconv/pytorch/layers/conv_chi_squared.py)."""
import numpy as np

from geoconv_b200.pytorch.layers.conv_exp import polar_distances, softmax_over_template
from geoconv_b200.pytorch.layers.conv_intrinsic import ConvIntrinsic


def gamma_func(dof):
    """The reference's recursion r * gamma_func(dof - 1) with gamma_func(1) = 1 (conv_chi_squared.py:8-28);
    kept as it is (it is not the Gamma function) so that the prior values agree."""
    assert dof >= 1, "You need to have at least one degree of freedom."
    r = dof / 2
    if dof == 1 / 2:
        return np.sqrt(np.pi)
    elif dof == 1:
        return 1.
    return r * gamma_func(dof - 1)


class ConvChiSquared(ConvIntrinsic):
    """Chi-squared vertex weighting (conv_chi_squared.py:31-94); with dof == 1 a zero radial or angular
    distance gets the weight 1 before the soft-max, as in the reference."""

    def __init__(self, *args, dof=2, **kwargs):
        self.dof = dof
        super().__init__(*args, **kwargs)

    def define_kernel_values(self, template_matrix):
        assert self.dof >= 1, "You need to have at least one degree of freedom."
        d_rho, d_theta = polar_distances(template_matrix)
        gamma = (1 / (2 ** (self.dof / 2) * gamma_func(self.dof))) ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            pdf = (gamma * d_rho ** (self.dof / 2 - 1) * d_theta ** (self.dof / 2 - 1)
                   * np.exp(-(d_rho + d_theta) / 2))
        if self.dof == 1:
            pdf = np.where((d_theta == 0) | (d_rho == 0), 1., pdf)
        return softmax_over_template(pdf)
