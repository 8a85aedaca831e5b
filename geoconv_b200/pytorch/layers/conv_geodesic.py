"""Drop-in for ``geoconv.pytorch.layers.conv_geodesic`` (/root/reference/src/geoconv/pytorch/layers/conv_geodesic.py)."""
import numpy as np

from geoconv_b200.pytorch.layers.conv_intrinsic import ConvIntrinsic


def angle_distance(theta_max, theta_min):
    """Shorter arc between two angles given theta_max >= theta_min (conv_geodesic.py:7-8)."""
    return np.minimum(theta_max - theta_min, theta_min + 2. * np.pi - theta_max)


def normal_pdf(mean_rho, mean_theta, var_rho, var_theta, rho, theta):
    """Axis-aligned Gaussian in geodesic polar coordinates with wrap-around angular distance
    (conv_geodesic.py:11-41).  Accepts scalars or broadcastable arrays."""
    d_theta = angle_distance(np.maximum(mean_theta, theta), np.minimum(mean_theta, theta))
    d_rho = rho - mean_rho
    norm = 1. / np.sqrt((2. * np.pi) ** 2 * var_rho * var_theta)
    return norm * np.exp(-.5 * (d_rho * d_rho / var_rho + d_theta * d_theta / var_theta))


class ConvGeodesic(ConvIntrinsic):
    """Geodesic convolution (Masci et al. 2015): Gaussian prior around every template vertex,
    soft-maxed over the template (conv_geodesic.py:44-68)."""

    def define_kernel_values(self, template_matrix):
        rho, theta = template_matrix[:, :, 0], template_matrix[:, :, 1]
        pdf = normal_pdf(rho[:, :, None, None], theta[:, :, None, None], rho.var(), theta.var(),
                         rho[None, None, :, :], theta[None, None, :, :])
        flat = pdf.reshape(pdf.shape[0], pdf.shape[1], -1)
        flat = np.exp(flat - flat.max(axis=-1, keepdims=True))
        return (flat / flat.sum(axis=-1, keepdims=True)).reshape(pdf.shape)
