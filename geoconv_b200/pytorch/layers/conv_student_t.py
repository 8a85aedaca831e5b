"""Drop-in for ``geoconv.pytorch.layers.conv_student_t``
(/root/reference/src/geoconv/pytorch/layers/conv_student_t.py)."""
import math

import numpy as np

from geoconv_b200.pytorch.layers.conv_exp import polar_distances, softmax_over_template
from geoconv_b200.pytorch.layers.conv_intrinsic import ConvIntrinsic


def gamma_func(x):
    """Gamma at integers and half-integers the way the reference evaluates it (conv_student_t.py:9-27;
    its ``np.math`` is the ``math`` module, which NumPy 2 no longer re-exports)."""
    assert x >= 1 / 2, "You need to have at least one degree of freedom."
    n = math.floor(x)
    if x - n == 1 / 2:
        return (math.factorial(2 * n) / (math.factorial(n) * 4 ** n)) * np.sqrt(np.pi)
    return math.factorial(n)


class ConvStudentT(ConvIntrinsic):
    """Student-t vertex weighting (conv_student_t.py:30-89)."""

    def __init__(self, *args, dof=2, **kwargs):
        self.dof = dof
        super().__init__(*args, **kwargs)

    def define_kernel_values(self, template_matrix):
        assert self.dof >= 1, "You need to have at least one degree of freedom."
        d_rho, d_theta = polar_distances(template_matrix)
        t_theta = (1 + d_theta ** 2 / self.dof) ** (-(self.dof + 1) / 2)
        t_rho = (1 + d_rho ** 2 / self.dof) ** (-(self.dof + 1) / 2)
        quotient = (gamma_func((self.dof + 1) / 2) / (np.sqrt(self.dof * np.pi) * gamma_func(self.dof / 2))) ** 2
        return softmax_over_template(quotient * t_rho * t_theta)
