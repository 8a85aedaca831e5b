"""Drop-in for ``geoconv.pytorch.layers.conv_exp`` (/root/reference/src/geoconv/pytorch/layers/conv_exp.py)."""
import numpy as np

from geoconv_b200.pytorch.layers.conv_geodesic import angle_distance
from geoconv_b200.pytorch.layers.conv_intrinsic import ConvIntrinsic


def softmax_over_template(pdf):
    """scipy.special.softmax over the whole (rho, theta) grid of every template vertex."""
    flat = pdf.reshape(pdf.shape[0], pdf.shape[1], -1)
    flat = np.exp(flat - flat.max(axis=-1, keepdims=True))
    return (flat / flat.sum(axis=-1, keepdims=True)).reshape(pdf.shape)


def polar_distances(template_matrix):
    """(|rho - mean_rho|, wrap-around |theta - mean_theta|) for every (mean, point) pair, after the radial
    coordinate has been normalised IN PLACE to [0, 1] exactly like the reference does (conv_exp.py:47)."""
    template_matrix[:, :, 0] = template_matrix[:, :, 0] / template_matrix[:, :, 0].max()
    rho, theta = template_matrix[:, :, 0], template_matrix[:, :, 1]
    d_rho = np.abs(rho[None, None, :, :] - rho[:, :, None, None])
    d_theta = angle_distance(np.maximum(theta[:, :, None, None], theta[None, None, :, :]),
                             np.minimum(theta[:, :, None, None], theta[None, None, :, :]))
    return d_rho, d_theta


def exp_pdf(mean_rho, mean_theta, rho, theta, exp_lambda):
    """Exponential weighting in geodesic polar coordinates (conv_exp.py:8-38)."""
    delta_angle = angle_distance(np.maximum(mean_theta, theta), np.minimum(mean_theta, theta))
    return exp_lambda ** 2 * np.exp(-exp_lambda * (np.abs(rho - mean_rho) + delta_angle))


class ConvExp(ConvIntrinsic):
    """Exponential vertex weighting (conv_exp.py:41-62)."""

    def __init__(self, *args, exp_lambda=1, **kwargs):
        self.exp_lambda = exp_lambda
        super().__init__(*args, **kwargs)

    def define_kernel_values(self, template_matrix):
        d_rho, d_theta = polar_distances(template_matrix)
        return softmax_over_template(self.exp_lambda ** 2 * np.exp(-self.exp_lambda * (d_rho + d_theta)))
