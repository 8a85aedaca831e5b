"""Drop-in for ``geoconv.pytorch.layers.angular_avg_pooling.AngularAvgPooling``
(/root/reference/src/geoconv/pytorch/layers/angular_avg_pooling.py:6-27)."""
import torch
from torch import nn

from geoconv_b200 import ops


class AngularAvgPooling(nn.Module):
    """Mean over the rotations: (V, n_rotations, T) -> (V, T)."""

    def forward(self, inputs):
        if inputs.dim() != 3:
            raise ValueError(f"expected (n_vertices, n_rotations, feature_dim), got {tuple(inputs.shape)}")
        z, _ = ops.AngularPoolFn.apply(inputs, ops.POOL_AVG)
        return z
