"""B200-native drop-in for ``geoconv.pytorch.layers.conv_intrinsic.ConvIntrinsic``
(/root/reference/src/geoconv/pytorch/layers/conv_intrinsic.py:32-262).

Same constructor arguments, same ``forward([signal, barycentric], orientations=None)`` call,
same parameter names / shapes / initialisation order (state_dict compatible), same
``define_kernel_values`` subclass hook and private attributes (``_kernel``, ``_template_size``,
``_template_vertices``).  The computation itself happens in libgeoconv_b200.so (hand-written
sm_100a CUDA, see include/geoconv_b200.h); this file only validates inputs and wires autograd.
CUDA only: calling the layer on CPU tensors raises.
"""
from abc import ABC, abstractmethod
import os

import numpy as np
import torch
from torch import nn

from geoconv_b200 import ops
from geoconv_b200._native import ACT_IDS
from geoconv_b200.preprocessing.barycentric_coordinates import create_template_matrix

# Same string keys as the reference tables (conv_intrinsic.py:10-29).  ACTIVATIONS values are the
# torch modules the fused epilogue reproduces; they are kept for API compatibility.
ACTIVATIONS = {
    "elu": nn.ELU(),
    "relu": nn.ReLU(),
    "leaky_relu": nn.LeakyReLU(),
    "selu": nn.SELU(),
    "sigmoid": nn.Sigmoid(),
    "tanh": nn.Tanh()
}

INITIALIZER = {
    "uniform": nn.init.uniform_,
    "normal": nn.init.normal_,
    "constant": nn.init.constant_,
    "xavier_uniform": nn.init.xavier_uniform_,
    "xavier_normal": nn.init.xavier_normal_,
    "kaiming_uniform": nn.init.kaiming_uniform_,
    "kaiming_normal": nn.init.kaiming_normal_,
    "trunc_normal": nn.init.trunc_normal_,
    "sparse_": nn.init.sparse_
}


def _rotation_symmetric(kernel, rotation_steps):
    """K[r,a,x,y] == K[r,(a+s)%A,x,(y+s)%A] for every used rotation step s (what makes it legal to
    fold the prior into the weights and express a rotation as an index permutation)."""
    for s in rotation_steps:
        if s and not np.allclose(kernel, np.roll(np.roll(kernel, s, axis=1), s, axis=3), rtol=0, atol=1e-7):
            return False
    return True


class ConvIntrinsic(ABC, nn.Module):
    """Intrinsic surface convolution on a triangle mesh (abstract; subclasses supply the prior).

    Attributes (identical to the reference)
    ----------
    activation_fn, rotation_delta, amt_templates, template_radius, initializer, include_prior

    Extra keyword (not in the reference)
    ----------
    precision: "fp32" (default; fp32-faithful, <=1e-5 normalised error: tensor cores on fp16 hi/lo operand
        splits - "f16x3" - where the shape allows, else on tf32 splits - "tf32x3" - else CUDA cores),
        "tf32" (<=2e-3) or "ffma" (force the CUDA-core kernels); "f16x3" / "tf32x3" select a split explicitly.
        Default can be overridden with GEOCONV_B200_PRECISION.
    """

    def __init__(self,
                 input_shape,
                 amt_templates,
                 template_radius,
                 include_prior=True,
                 activation="relu",
                 rotation_delta=1,
                 initializer="xavier_uniform",
                 precision=None):
        super().__init__()
        self.activation_fn = activation
        self.rotation_delta = rotation_delta
        self.amt_templates = amt_templates
        self.template_radius = template_radius
        self.initializer = initializer
        self.include_prior = include_prior
        self.precision = precision or os.environ.get("GEOCONV_B200_PRECISION", "fp32")
        self.backward_precision = None  # None: same as `precision`
        self.check_indices = True

        self._activation = ACTIVATIONS[self.activation_fn]
        self._act_id = ACT_IDS[self.activation_fn]
        self._init_fn = INITIALIZER[self.initializer]
        self._bias = None
        self._all_rotations = None
        self._template_size = None  # (#radial, #angular)
        self._template_neighbor_weights = None
        self._template_self_weights = None
        self._feature_dim = None
        self._prior_is_zero = False

        self.build(input_shape)

    def build(self, input_shape):
        """Creates template, parameters (in the reference's order, conv_intrinsic.py:85-116, so a
        seeded construction yields identical initial weights) and the prior."""
        signal_shape, barycentric_shape = input_shape

        self._template_size = (barycentric_shape[1], barycentric_shape[2])
        self._all_rotations = self._template_size[1]
        # non-persistent buffers: follow .to(device) but stay out of the state_dict, like the
        # reference's plain tensor attributes
        self.register_buffer(
            "_template_vertices",
            torch.tensor(create_template_matrix(self._template_size[0], self._template_size[1],
                                                radius=self.template_radius)),
            persistent=False)
        self._feature_dim = signal_shape[-1]

        self._template_neighbor_weights = nn.Parameter(
            torch.zeros(size=(self.amt_templates, self._template_size[0], self._template_size[1], signal_shape[1]))
        )
        self._init_fn(self._template_neighbor_weights)

        self._template_self_weights = nn.Parameter(torch.zeros(size=(self.amt_templates, 1, signal_shape[1])))
        self._init_fn(self._template_self_weights)

        self._bias = nn.Parameter(torch.zeros(size=(1, self.amt_templates)))
        self._init_fn(self._bias)

        self._configure_kernel()

    def _configure_kernel(self):
        """Evaluates the subclass prior once (conv_intrinsic.py:242-244) and classifies it."""
        values = np.asarray(self.define_kernel_values(self._template_vertices.cpu().numpy())).astype(np.float32)
        self.register_buffer("_kernel", torch.tensor(values), persistent=False)
        self._kernel_host = values          # numpy copy for host-side checks (no device sync per step)
        self._symmetry_checked = {}
        self._prior_is_zero = bool(self.include_prior and not np.any(values))
        steps = range(0, self._all_rotations, max(int(self.rotation_delta), 1))
        self._prior_is_symmetric = _rotation_symmetric(values, steps) if self.include_prior else True
        # the layer's own rotation set is answered here, so a traced forward only looks the result up
        self._symmetry_checked[tuple(sorted(set(steps)))] = self._prior_is_symmetric

    def _rotation_offsets(self, orientations):
        if orientations is None:
            return list(range(0, self._all_rotations, self.rotation_delta))
        if torch.is_tensor(orientations):
            orientations = orientations.detach().cpu().tolist()
        return [int(o) for o in orientations]

    def _effective_weights(self, offsets):
        if not self.include_prior:
            return self._template_neighbor_weights
        if self._prior_is_zero:
            return None
        steps = tuple(sorted({o % self._all_rotations for o in offsets}))
        if steps not in self._symmetry_checked:   # host-side check once per rotation set, never per step
            self._symmetry_checked[steps] = _rotation_symmetric(self._kernel_host, steps)
        if not self._symmetry_checked[steps]:
            raise NotImplementedError(
                "geoconv_b200: the prior returned by define_kernel_values is not invariant under the requested "
                "template rotations (K[r,a,x,y] != K[r,a+s,x,y+s]); only rotation-symmetric priors (every prior "
                "shipped by geoconv: functions of the radial and angular distance) are supported.")
        return ops.FoldPrior.apply(self._template_neighbor_weights, self._kernel)

    def forward(self, inputs, orientations=None, row_offset=0):
        """inputs = [signal (V, F), barycentric (V, R, A, 3, 2)] -> (V, n_rotations, templates).

        ``row_offset`` (extension, forward only): the barycentric tensor holds the rows of vertices
        ``row_offset .. row_offset + len(barycentric)`` of the signal (its indices still address the whole signal).

        The result also carries the fused AngularMaxPooling outputs so that a following
        ``AngularMaxPooling`` module returns them without re-reading the tensor and back-propagates
        through one rotation per vertex only.
        """
        y, z, arg = self._run(inputs, orientations, row_offset)
        if torch.compiler.is_compiling():
            # Dynamo tracks attributes only on tensors made by a traced torch call (an operator's tuple output is
            # not one), and cannot read the version counter: inside a trace the tuple rides on an alias of Y and
            # lives and dies with that trace
            y = y.view_as(y)
            y._gcb_pooled = (z, arg, None)
        else:
            y._gcb_pooled = (z, arg, y._version)
        return y

    def forward_pooled(self, inputs, orientations=None):
        """Fused convolution + angular max-pooling: returns (Z (V, T), argmax (V,) int32)."""
        _, z, arg = self._run(inputs, orientations)
        return z, arg

    def _run(self, inputs, orientations, row_offset=0):
        # Traceable by torch.compile (the reference demo compiles its model, training_demo.py:126-127): every
        # device-side step is a `geoconv_b200::*` torch.library operator with a fake kernel, the rest is shape logic.
        mesh_signal, bary_coordinates = inputs
        ops._need_cuda(mesh_signal, "signal")
        ops._need_cuda(bary_coordinates, "barycentric coordinates")
        ops._need_cuda(self._template_self_weights, "layer parameters")
        if mesh_signal.dim() != 2 or mesh_signal.shape[1] != self._feature_dim:
            raise ValueError(f"signal must have shape (V, {self._feature_dim}), got {tuple(mesh_signal.shape)}")
        if tuple(bary_coordinates.shape[1:3]) != tuple(self._template_size):
            raise ValueError(f"barycentric coordinates must have shape (V, {self._template_size[0]}, "
                             f"{self._template_size[1]}, 3, 2), got {tuple(bary_coordinates.shape)}")
        if row_offset < 0 or row_offset + bary_coordinates.shape[0] > mesh_signal.shape[0]:
            raise ValueError("more barycentric rows than signal rows")
        signal = mesh_signal.to(torch.float32)
        if torch.compiler.is_compiling():
            # no Python-side cache inside a traced graph: the split is one more operator of the graph
            pack = ops.prep_bary_traced(bary_coordinates, signal.shape[0])
        else:
            ops.INDEX_CHECKS.poll()       # deferred index validation of earlier tensors (no host sync)
            pack = ops.GLOBAL_BARY_CACHE.get(bary_coordinates, signal.shape[0], self.check_indices)
        offsets = self._rotation_offsets(orientations)
        w_eff = self._effective_weights(offsets)
        r, a = self._template_size
        prec = ops.resolve_precision(self.precision, r, a, self._feature_dim, self.amt_templates, len(offsets))
        if self.backward_precision is not None:
            prec = (prec, ops.resolve_precision(self.backward_precision, r, a, self._feature_dim,
                                                self.amt_templates, len(offsets)))
        anchor = self._template_neighbor_weights if w_eff is None else None
        return ops.ConvIntrinsicFn.apply(signal, w_eff, self._template_self_weights, self._bias, pack,
                                         tuple(offsets), self._act_id, prec, anchor, row_offset)

    @abstractmethod
    def define_kernel_values(self, template_matrix):
        """Interpolation coefficients of the patch operator.

        template_matrix: np.ndarray (n_radial, n_angular, 2) of polar template coordinates.
        Returns np.ndarray (n_radial, n_angular, n_radial, n_angular).
        """
        pass
