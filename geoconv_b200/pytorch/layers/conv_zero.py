"""Drop-in for ``geoconv.pytorch.layers.conv_zero`` (/root/reference/src/geoconv/pytorch/layers/conv_zero.py)."""
import numpy as np

from geoconv_b200.pytorch.layers.conv_intrinsic import ConvIntrinsic


class ConvZero(ConvIntrinsic):
    """All-zero prior: only the self-connection contributes (conv_zero.py:6-15).  The kernels skip
    the neighbour contraction entirely; the neighbour-weight gradient is an all-zeros tensor."""

    def define_kernel_values(self, template_matrix):
        return np.zeros(template_matrix.shape[:-1] + template_matrix.shape[:-1])
