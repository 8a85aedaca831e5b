"""Drop-in for ``geoconv.pytorch.layers.conv_dirac`` (/root/reference/src/geoconv/pytorch/layers/conv_dirac.py)."""
import numpy as np

from geoconv_b200.pytorch.layers.conv_intrinsic import ConvIntrinsic


class ConvDirac(ConvIntrinsic):
    """Dirac prior: interpolated signals are used as they are, so the patch operator is skipped
    (``include_prior`` is forced to False, conv_dirac.py:9-11)."""

    def __init__(self, *args, **kwargs):
        kwargs["include_prior"] = False
        super().__init__(*args, **kwargs)

    def define_kernel_values(self, template_matrix):
        """Identity coefficients; only used for visualisation (conv_dirac.py:13-26)."""
        r, a = template_matrix.shape[:2]
        return np.eye(r * a).reshape(r, a, r, a)
