"""``BatchNormDropout``: the chain the reference's network runs BETWEEN two convolutions, as one module
(SURVEY.md §8 f-2).  /root/reference/src/geoconv_examples/mpi_faust/pytorch/model.py:93-105 builds, per block,
``nn.Dropout(p=0.2)``, the ISC layer, ``AngularMaxPooling`` and ``nn.BatchNorm1d``; ``forward`` (:128-132) runs
``do -> isc -> amp -> bn``, so the batch norm of block l feeds the dropout of block l + 1.  This module is that
``bn -> do`` pair: statistics (``gcb_bn_stats``), then normalisation, affine map and dropout in ONE pass over the
(V, T) tensor (``gcb_bn_dropout_fwd``), two passes in the backward (``gcb_bn_dropout_bwd``) - instead of the ~9
kernels PyTorch launches for the two modules.  CUDA only.

Parameters and buffers carry ``nn.BatchNorm1d``'s names (``weight``, ``bias``, ``running_mean``, ``running_var``,
``num_batches_tracked``), so a reference checkpoint's ``bn_layers.<i>.*`` entries load into it unchanged.
The dropout mask is a function of (seed, call counter, element index) (Philox 4x32-10) and is never stored; with
``p = 0`` or in ``eval()`` the module is exactly BatchNorm1d.  Replaying a captured CUDA graph replays its masks.
"""
import torch
from torch import nn

from geoconv_b200 import ops

_CALLS = 0


class BatchNormDropout(nn.Module):
    def __init__(self, num_features, p=0.2, eps=1e-5, momentum=0.1):
        super().__init__()
        if num_features % 4:
            raise ValueError("geoconv_b200: num_features must be a multiple of 4")
        if not 0.0 <= p < 1.0:
            raise ValueError("dropout probability has to be in [0, 1)")
        self.num_features, self.p, self.eps, self.momentum = num_features, float(p), float(eps), momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self._steps = 0      # host mirror of num_batches_tracked (cumulative average without a device read)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        key = prefix + "num_batches_tracked"
        if key in state_dict:
            self._steps = int(state_dict[key])

    def forward(self, z):
        global _CALLS
        if not z.is_cuda:
            raise RuntimeError("geoconv_b200: BatchNormDropout runs on a CUDA device (no CPU path)")
        if z.dim() != 2 or z.shape[1] != self.num_features:
            raise ValueError(f"expected (V, {self.num_features}), got {tuple(z.shape)}")
        z = z.to(torch.float32)
        if self.training:
            self._steps += 1
            with torch.no_grad():
                self.num_batches_tracked += 1
            m = self.momentum if self.momentum is not None else 1.0 / self._steps
            mean, invstd = ops.bn_stats_op(z.detach(), self.eps, self.running_mean, self.running_var, m)
            _CALLS += 1
            return ops.bn_dropout_op(z, mean, invstd, self.weight, self.bias, True, self.p,
                                     torch.initial_seed() & (2 ** 63 - 1), _CALLS)
        invstd = torch.rsqrt(self.running_var + self.eps)
        return ops.bn_dropout_op(z, self.running_mean, invstd, self.weight, self.bias, False, 0.0, 0, 0)

    def extra_repr(self):
        return f"{self.num_features}, p={self.p}, eps={self.eps}, momentum={self.momentum}"
