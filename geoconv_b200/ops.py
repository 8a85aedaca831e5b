"""Torch-facing wrappers over the C ABI: device-memory plumbing and autograd wiring only.

Every function here hands raw device pointers of contiguous CUDA tensors to libgeoconv_b200.so on
the current stream.  Nothing in this file computes the convolution with torch ops.
"""
from __future__ import annotations

import os
import weakref
from dataclasses import dataclass, field

import torch

from . import _native as N

# Keep the interpolations of a training forward for its backward (V*R*A*F floats per call).
# Measured on B200 at the FAUST shape this is SLOWER than gathering again (dW 422 vs 342 us, forward
# +90 us): the 600 MB tensor streams from HBM in scattered 128-byte pieces while the 15 MB signal
# it is built from stays in L2.  Off by default; the switch remains for shapes with a tiny signal.
SAVE_INTERPOLATIONS = False

# 'fp32' (the fp32-faithful default) runs the tensor-core kernels with fp16 hi/lo operand splits (three
# kind::f16 products: same significand bits as 3xTF32 at twice the rate) when the shape allows;
# GEOCONV_B200_FP32_VIA=tf32x3 selects the tf32 splits instead.
FP32_VIA_F16 = os.environ.get("GEOCONV_B200_FP32_VIA", "f16x3") != "tf32x3"

PRECISIONS = {"fp32": None, "ffma": N.PREC_FP32, "tf32": N.PREC_TF32, "tf32x3": N.PREC_TF32X3,
              "f16x3": N.PREC_F16X3}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"geoconv_b200: `{name}` lives on {t.device}; this implementation is CUDA (sm_100a) only "
            "and has no CPU path - move the module and its inputs to a B200 device.")


def resolve_precision(mode: str, r: int, a: int, f: int, t: int, n_rot: int) -> int:
    """'fp32' = fp32-faithful: tensor cores on fp16 hi/lo splits (F % 16 == 0), else on tf32 splits (3xTF32),
    else CUDA cores; explicit tensor modes fall back to the exact CUDA-core arithmetic for shapes they cannot tile."""
    if mode not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {mode!r}")
    lib = N.load()
    tc_ok = bool(lib.gcb_tc_supported(r, a, f, t, n_rot))
    half_ok = tc_ok and bool(lib.gcb_tc_half_supported(r, a, f, t, n_rot))
    if mode == "fp32":
        if not tc_ok:
            return N.PREC_FP32
        return N.PREC_F16X3 if half_ok and FP32_VIA_F16 else N.PREC_TF32X3
    code = PRECISIONS[mode]
    if code == N.PREC_F16X3 and not half_ok:
        code = N.PREC_TF32X3
    if code != N.PREC_FP32 and not tc_ok:
        return N.PREC_FP32  # exact arithmetic is always an admissible substitute for a faster mode
    return code


# ------------------------------------------------------------------------------------------------
# bary -> packed (idx, w) [+ lazily the CSR inversion], cached per bary tensor object
# ------------------------------------------------------------------------------------------------
@dataclass
class BaryPack:
    idx: torch.Tensor          # int32 [V_out,R,A,3]
    w: torch.Tensor            # fp32  [V_out,R,A,3]
    v_out: int
    v_src: int
    r: int
    a: int
    _csr: tuple | None = field(default=None, repr=False)

    def csr(self):
        """(row_ptr int32 [V_src+1], entries int32 [n_slots]) - built on first backward."""
        if self._csr is None:
            lib = N.load()
            n_slots = self.idx.numel()
            dev = self.idx.device
            row_ptr = torch.empty(self.v_src + 1, dtype=torch.int32, device=dev)
            entries = torch.empty(n_slots, dtype=torch.int32, device=dev)
            ws_bytes = lib.gcb_csr_workspace_bytes(n_slots, self.v_src)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                N.check(lib.gcb_csr_build(self.idx.data_ptr(), self.w.data_ptr(), n_slots, self.v_src,
                                          row_ptr.data_ptr(), entries.data_ptr(), ws.data_ptr(), ws_bytes,
                                          _stream(self.idx)), "gcb_csr_build")
            self._csr = (row_ptr, entries)
        return self._csr


def prep_bary(bary: torch.Tensor, v_src: int, check_indices: bool = True) -> BaryPack:
    """Split a (V,R,A,3,2) barycentric tensor (fp32 or fp64) into int32 indices and fp32 weights.

    Same arithmetic as the reference: cast to fp32, truncate the index column toward zero.
    With ``check_indices`` an out-of-range index raises IndexError (one host sync per *new* bary
    tensor; the reference's CPU gather raises too, its CUDA gather is undefined behaviour).
    """
    _need_cuda(bary, "barycentric coordinates")
    if bary.dim() != 5 or bary.shape[3] != 3 or bary.shape[4] != 2:
        raise ValueError(f"barycentric coordinates must have shape (V,R,A,3,2), got {tuple(bary.shape)}")
    if bary.dtype not in (torch.float32, torch.float64):
        bary = bary.to(torch.float32)
    bary = bary.contiguous()
    v, r, a = bary.shape[:3]
    lib = N.load()
    idx = torch.empty((v, r, a, 3), dtype=torch.int32, device=bary.device)
    w = torch.empty((v, r, a, 3), dtype=torch.float32, device=bary.device)
    oob = torch.zeros(1, dtype=torch.int32, device=bary.device)
    if idx.numel():
        with torch.cuda.device(bary.device):
            N.check(lib.gcb_prep_bary(bary.data_ptr(), int(bary.dtype == torch.float64), idx.numel(), v_src,
                                      idx.data_ptr(), w.data_ptr(), oob.data_ptr(), _stream(bary)),
                    "gcb_prep_bary")
        if check_indices:
            bad = int(oob.item())
            if bad:
                raise IndexError(f"geoconv_b200: {bad} barycentric vertex indices fall outside [0, {v_src})")
    return BaryPack(idx=idx, w=w, v_out=v, v_src=v_src, r=r, a=a)


class BaryCache:
    """Per-layer-instance-independent cache: bary tensor object -> BaryPack.

    Keyed on the tensor *object* (weak reference) and its version counter, never on data_ptr
    alone: the caching allocator recycles addresses."""

    def __init__(self, capacity: int = 16):
        self.capacity = capacity
        self._items: dict[int, tuple] = {}

    def get(self, bary: torch.Tensor, v_src: int, check_indices: bool = True) -> BaryPack:
        key = id(bary)
        hit = self._items.get(key)
        if hit is not None:
            ref, version, ptr, vs, pack = hit
            if ref() is bary and version == bary._version and ptr == bary.data_ptr() and vs == v_src:
                return pack
        pack = prep_bary(bary, v_src, check_indices)
        if len(self._items) >= self.capacity:
            dead = [k for k, (ref, *_rest) in self._items.items() if ref() is None]
            for k in dead or list(self._items)[:1]:
                self._items.pop(k, None)
        self._items[key] = (weakref.ref(bary), bary._version, bary.data_ptr(), v_src, pack)
        return pack


GLOBAL_BARY_CACHE = BaryCache()


# ------------------------------------------------------------------------------------------------
# prior folding  W[T,R,A,F] x K[R,A,R,A] -> Weff
# ------------------------------------------------------------------------------------------------
def _fold(inp: torch.Tensor, prior: torch.Tensor, transpose: int) -> torch.Tensor:
    lib = N.load()
    t, r, a, f = inp.shape
    out = torch.empty_like(inp)
    with torch.cuda.device(inp.device):
        N.check(lib.gcb_fold_prior(inp.data_ptr(), prior.data_ptr(), out.data_ptr(), t, r, a, f, transpose,
                                   _stream(inp)), "gcb_fold_prior")
    return out


class FoldPrior(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w_neighbor, prior):
        _need_cuda(w_neighbor, "template weights")
        prior = prior.to(device=w_neighbor.device, dtype=torch.float32).contiguous()
        ctx.save_for_backward(prior)
        return _fold(w_neighbor.contiguous(), prior, 0)

    @staticmethod
    def backward(ctx, grad):
        (prior,) = ctx.saved_tensors
        return _fold(grad.contiguous(), prior, 1), None


# ------------------------------------------------------------------------------------------------
# the convolution (+ fused pooling outputs)
# ------------------------------------------------------------------------------------------------
class ConvIntrinsicFn(torch.autograd.Function):
    """(signal, Weff|None, Wself, bias) -> (Y [V,n_rot,T], Z [V,T], argmax [V]).

    Y is the layer output; Z/argmax are what AngularMaxPooling would produce from it.  In backward,
    a gradient arriving only through Z takes the AMP-sparse route (one rotation per vertex); a
    gradient arriving through Y is back-propagated one rotation at a time.
    """

    @staticmethod
    def forward(ctx, signal, w_eff, w_self, bias, pack: BaryPack, rot_offsets, act_id, precision, w_anchor=None):
        lib = N.load()
        _need_cuda(signal, "signal")
        # `precision` is one code for both directions or a (forward, backward) pair
        precision, precision_bwd = precision if isinstance(precision, tuple) else (precision, precision)
        x = signal.contiguous()
        v_src, f = x.shape
        t = w_self.shape[0]
        n_rot = len(rot_offsets)
        dev = x.device
        w_self_c = w_self.reshape(t, f).contiguous()
        bias_c = bias.reshape(t).contiguous()
        w_eff_c = None if w_eff is None else w_eff.contiguous()
        y = torch.empty((pack.v_out, n_rot, t), dtype=torch.float32, device=dev)
        z = torch.empty((pack.v_out, t), dtype=torch.float32, device=dev)
        arg = torch.empty((pack.v_out,), dtype=torch.int32, device=dev)
        rot = N.rot_array(rot_offsets)
        # Training: let the tensor-core forward keep the interpolations I [V,R,A,F] (the reference
        # materialises the same tensor) so the backward's dW contraction reads them back instead of
        # gathering three signal rows per template vertex again.
        keep = (SAVE_INTERPOLATIONS and w_eff_c is not None and precision != N.PREC_FP32
                and precision_bwd != N.PREC_FP32 and any(ctx.needs_input_grad[:4]))
        interp = torch.empty((pack.v_out, pack.r * pack.a * f), dtype=torch.float32, device=dev) if keep else None
        ws_bytes = lib.gcb_conv_fwd_workspace_bytes(pack.v_out, v_src, pack.r, pack.a, f, t, n_rot, precision)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            N.check(lib.gcb_conv_fwd(x.data_ptr(), pack.idx.data_ptr(), pack.w.data_ptr(), _ptr(w_eff_c),
                                     w_self_c.data_ptr(), bias_c.data_ptr(), rot, n_rot, pack.v_out, v_src,
                                     pack.r, pack.a, f, t, act_id, precision, y.data_ptr(), z.data_ptr(),
                                     arg.data_ptr(), _ptr(interp), ws.data_ptr(), ws_bytes, _stream(x)),
                    "gcb_conv_fwd")
        ctx.pack = pack
        ctx.rot_offsets = tuple(int(o) for o in rot_offsets)
        ctx.act_id = act_id
        ctx.precision = precision_bwd
        ctx.has_neighbor = w_eff is not None
        # all-zero prior: the neighbour weights still receive an all-zeros gradient, like the reference
        ctx.anchor_shape = None if w_anchor is None else tuple(w_anchor.shape)
        ctx.w_self_shape = tuple(w_self.shape)
        ctx.bias_shape = tuple(bias.shape)
        ctx.has_interp = interp is not None
        ctx.save_for_backward(x, w_eff_c if w_eff_c is not None else x.new_empty(0), w_self_c, y, z, arg,
                              interp if interp is not None else x.new_empty(0))
        ctx.mark_non_differentiable(arg)
        ctx.set_materialize_grads(False)
        return y, z, arg

    @staticmethod
    def backward(ctx, grad_y, grad_z, _grad_arg):
        lib = N.load()
        x, w_eff, w_self, y, z, arg, interp = ctx.saved_tensors
        interp_ptr = interp.data_ptr() if ctx.has_interp else None
        pack: BaryPack = ctx.pack
        has_nb = ctx.has_neighbor
        v_src, f = x.shape
        t = w_self.shape[0]
        dev = x.device
        n_rot = len(ctx.rot_offsets)
        need_dx = ctx.needs_input_grad[0]
        if grad_y is None and grad_z is None:
            return (None,) * 9
        d_x = torch.empty_like(x) if need_dx else None
        d_w_eff = torch.empty((t, pack.r, pack.a, f), dtype=torch.float32, device=dev)
        d_w_self = torch.empty((t, f), dtype=torch.float32, device=dev)
        d_bias = torch.empty((t,), dtype=torch.float32, device=dev)
        ws_bytes = lib.gcb_conv_bwd_workspace_bytes(pack.v_out, v_src, pack.r, pack.a, f, t, ctx.precision)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        # the CSR inversion also feeds the weight gradients (dW = G^T X on tensor cores), not only dSignal
        row_ptr, entries = pack.csr() if has_nb else (None, None)
        h = torch.empty((pack.v_out, t), dtype=torch.float32, device=dev)
        rot_sel = torch.empty((pack.v_out,), dtype=torch.int32, device=dev)
        st = _stream(x)
        rot = N.rot_array(ctx.rot_offsets)
        calls = 0

        def run(accumulate):
            N.check(lib.gcb_conv_bwd(x.data_ptr(), pack.idx.data_ptr(), pack.w.data_ptr(),
                                     w_eff.data_ptr() if has_nb else None, w_self.data_ptr(), h.data_ptr(),
                                     rot_sel.data_ptr(), _ptr(row_ptr), _ptr(entries), pack.v_out, v_src,
                                     pack.r, pack.a, f, t, ctx.precision, accumulate, interp_ptr, _ptr(d_x),
                                     d_w_eff.data_ptr(), d_w_self.data_ptr(), d_bias.data_ptr(),
                                     ws.data_ptr(), ws_bytes, st), "gcb_conv_bwd")

        with torch.cuda.device(dev):
            if grad_z is not None:
                gz = grad_z.contiguous().to(torch.float32)
                N.check(lib.gcb_act_grad(gz.data_ptr(), z.data_ptr(), gz.numel(), ctx.act_id, h.data_ptr(), st),
                        "gcb_act_grad")
                N.check(lib.gcb_select_rot(arg.data_ptr(), rot, n_rot, pack.v_out, rot_sel.data_ptr(), st),
                        "gcb_select_rot")
                run(0)
                calls += 1
            if grad_y is not None:
                gy = grad_y.to(torch.float32)
                for i, off in enumerate(ctx.rot_offsets):
                    gi = gy[:, i, :].contiguous()
                    yi = y[:, i, :].contiguous()
                    N.check(lib.gcb_act_grad(gi.data_ptr(), yi.data_ptr(), gi.numel(), ctx.act_id, h.data_ptr(),
                                             st), "gcb_act_grad")
                    rot_sel.fill_(off % pack.a)
                    run(1 if calls else 0)
                    calls += 1
        g_w_eff = d_w_eff if has_nb else None
        g_anchor = None if ctx.anchor_shape is None else torch.zeros(ctx.anchor_shape, dtype=torch.float32, device=dev)
        return (d_x, g_w_eff, d_w_self.reshape(ctx.w_self_shape), d_bias.reshape(ctx.bias_shape),
                None, None, None, None, g_anchor)


# ------------------------------------------------------------------------------------------------
# stand-alone AngularMaxPooling on a materialised tensor
# ------------------------------------------------------------------------------------------------
class AngularMaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y):
        lib = N.load()
        _need_cuda(y, "inputs")
        yc = y.contiguous().to(torch.float32)
        v, n_rot, t = yc.shape
        z = torch.empty((v, t), dtype=torch.float32, device=yc.device)
        arg = torch.empty((v,), dtype=torch.int32, device=yc.device)
        with torch.cuda.device(yc.device):
            N.check(lib.gcb_amp_fwd(yc.data_ptr(), v, n_rot, t, z.data_ptr(), arg.data_ptr(), _stream(yc)),
                    "gcb_amp_fwd")
        ctx.save_for_backward(arg)
        ctx.shape = (v, n_rot, t)
        ctx.mark_non_differentiable(arg)
        return z, arg

    @staticmethod
    def backward(ctx, grad_z, _):
        lib = N.load()
        (arg,) = ctx.saved_tensors
        v, n_rot, t = ctx.shape
        gz = grad_z.contiguous().to(torch.float32)
        gy = torch.empty((v, n_rot, t), dtype=torch.float32, device=gz.device)
        with torch.cuda.device(gz.device):
            N.check(lib.gcb_amp_bwd(gz.data_ptr(), arg.data_ptr(), v, n_rot, t, gy.data_ptr(), _stream(gz)),
                    "gcb_amp_bwd")
        return gy


POOL_MAX, POOL_MIN, POOL_AVG = 0, 1, 2


class AngularPoolFn(torch.autograd.Function):
    """Angular max / min / average pooling of a materialised (V, n_rot, T) tensor."""

    @staticmethod
    def forward(ctx, y, mode):
        lib = N.load()
        _need_cuda(y, "inputs")
        yc = y.contiguous().to(torch.float32)
        v, n_rot, t = yc.shape
        z = torch.empty((v, t), dtype=torch.float32, device=yc.device)
        arg = torch.empty((v,), dtype=torch.int32, device=yc.device)
        with torch.cuda.device(yc.device):
            N.check(lib.gcb_apool_fwd(yc.data_ptr(), v, n_rot, t, mode, z.data_ptr(), arg.data_ptr(), _stream(yc)),
                    "gcb_apool_fwd")
        ctx.save_for_backward(arg)
        ctx.shape = (v, n_rot, t)
        ctx.mode = mode
        ctx.mark_non_differentiable(arg)
        return z, arg

    @staticmethod
    def backward(ctx, grad_z, _):
        lib = N.load()
        (arg,) = ctx.saved_tensors
        v, n_rot, t = ctx.shape
        gz = grad_z.contiguous().to(torch.float32)
        gy = torch.empty((v, n_rot, t), dtype=torch.float32, device=gz.device)
        with torch.cuda.device(gz.device):
            N.check(lib.gcb_apool_bwd(gz.data_ptr(), arg.data_ptr(), v, n_rot, t, ctx.mode, gy.data_ptr(),
                                      _stream(gz)), "gcb_apool_bwd")
        return gy, None


def pack_rows(src: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """out[i] = src[rows[i]] (halo send-buffer packing)."""
    lib = N.load()
    _need_cuda(src, "src")
    src = src.contiguous()
    rows = rows.to(device=src.device, dtype=torch.int32).contiguous()
    out = torch.empty((rows.numel(), src.shape[1]), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        N.check(lib.gcb_pack_rows(src.data_ptr(), rows.data_ptr(), rows.numel(), src.shape[1], out.data_ptr(),
                                  _stream(src)), "gcb_pack_rows")
    return out
