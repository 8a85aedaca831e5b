"""Drop-in for ``geoconv.pytorch.layers.angular_min_pooling.AngularMinPooling``
(/root/reference/src/geoconv/pytorch/layers/angular_min_pooling.py:6-29)."""
import torch
from torch import nn

from geoconv_b200 import ops


class AngularMinPooling(nn.Module):
    """Per vertex, keep the rotation whose template-response vector has the SMALLEST Euclidean norm
    (first minimum wins): (V, n_rotations, T) -> (V, T)."""

    def forward(self, inputs):
        if inputs.dim() != 3:
            raise ValueError(f"expected (n_vertices, n_rotations, feature_dim), got {tuple(inputs.shape)}")
        z, _ = ops.AngularPoolFn.apply(inputs, ops.POOL_MIN)
        return z


# The reference's angular_min_pooling.py names its class ``AngularMaxPooling`` (angular_min_pooling.py:6);
# code written against it imports that name from this module path.
AngularMaxPooling = AngularMinPooling
