"""Torch-facing wrappers over the C ABI: device-memory plumbing and autograd wiring only.

Every function here hands raw device pointers of contiguous CUDA tensors to libgeoconv_b200.so on
the current stream.  Nothing in this file computes the convolution with torch ops.
"""
from __future__ import annotations

import contextlib
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
    """Raw handle of the current stream of t's device (the Python Stream object costs ~10 us per look-up)."""
    idx = t.device.index
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device() if idx is None else idx)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"geoconv_b200: `{name}` lives on {t.device}; this implementation is CUDA (sm_100a) only "
            "and has no CPU path - move the module and its inputs to a B200 device.")


@torch.compiler.assume_constant_result
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
class _IndexChecks:
    """Deferred validation of barycentric vertex indices.

    ``gcb_prep_bary`` counts the indices outside [0, V_src) on the device (and neutralises them, so the
    kernels never read out of bounds).  Instead of reading that counter back with a blocking ``.item()``
    per new barycentric tensor, the counter is copied to a pinned host slot behind an event; every later
    layer call / backward polls the finished events (no host sync) and raises IndexError for a tensor that
    held bad indices - the same deferred reporting CUDA itself uses for asynchronous errors.
    ``BaryPack.validate()`` blocks until the answer is known."""

    SLOTS = 512

    def __init__(self):
        self._host = None
        self._next = 0
        self._pending: list[tuple] = []   # (event, slot, v_src, weakref to the pack)

    def submit(self, pack: "BaryPack", oob: torch.Tensor) -> None:
        if torch.cuda.is_current_stream_capturing():
            return                       # a captured step cannot report back to the host
        if self._host is None:
            self._host = torch.zeros(self.SLOTS, dtype=torch.int32).pin_memory()
        if len(self._pending) >= self.SLOTS:
            self.poll(block=True)
        slot = self._next
        self._next = (self._next + 1) % self.SLOTS
        self._host[slot:slot + 1].copy_(oob, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(oob.device))
        self._pending.append((ev, slot, pack.v_src, weakref.ref(pack)))

    def poll(self, block: bool = False, only: "BaryPack | None" = None) -> None:
        if not self._pending or torch.cuda.is_current_stream_capturing():
            return                       # (an event query is not allowed while a stream is being captured)
        keep, bad = [], None
        for item in self._pending:
            ev, slot, v_src, ref = item
            mine = only is None or ref() is only
            if mine and block:
                ev.synchronize()
            if mine and ev.query():
                n = int(self._host[slot])
                pk = ref()
                if pk is not None:
                    pk.oob_count = n
                if n and bad is None:
                    bad = (n, v_src)
            else:
                keep.append(item)
        self._pending = keep
        if bad is not None:
            raise IndexError(f"geoconv_b200: {bad[0]} barycentric vertex indices fall outside [0, {bad[1]})")


INDEX_CHECKS = _IndexChecks()


@dataclass(eq=False)
class BaryPack:
    idx: torch.Tensor          # int32 [V_out,R,A,3]
    w: torch.Tensor            # fp32  [V_out,R,A,3]
    v_out: int
    v_src: int
    r: int
    a: int
    oob_count: int | None = None          # out-of-range indices found by gcb_prep_bary (None: not read back yet)
    keeps_all_slots: bool = False         # the CSR inversion keeps zero-weight slots (shared between weight variants)
    _csr: tuple | None = field(default=None, repr=False)
    _records: torch.Tensor | None = field(default=None, repr=False)

    def validate(self) -> None:
        """Block until the index check of this tensor is known; raises IndexError if it failed."""
        INDEX_CHECKS.poll(block=True, only=self)
        if self.oob_count:
            raise IndexError(f"geoconv_b200: {self.oob_count} barycentric vertex indices fall outside [0, {self.v_src})")

    def csr(self):
        """(row_ptr int32 [V_src+1], entries int32 [n_slots], stats int32 [8]) - built on first backward."""
        if self._csr is None:
            # (all-ones weights: nothing is dropped and the stats bound |w| by 1, which barycentric weights respect)
            w = torch.ones_like(self.w) if self.keeps_all_slots else self.w
            self._csr = tuple(csr_build_op(self.idx, w, self.v_src))
        return self._csr

    def records(self):
        """The tensor-core forward's packed gather records of this tensor (built on first use, then reused by
        every layer and every step that sees the same barycentric tensor)."""
        if self._records is None:
            self._records = pack_records_op(self.idx, self.w)
        return self._records


def prep_bary(bary: torch.Tensor, v_src: int, check_indices: bool | str = True,
              keep_zero_weight_slots: bool = False) -> BaryPack:
    """Split a (V,R,A,3,2) barycentric tensor (fp32 or fp64) into int32 indices and fp32 weights.

    Same arithmetic as the reference: cast to fp32, truncate the index column toward zero.
    Out-of-range indices (the reference's CPU gather raises on them, its CUDA gather is undefined
    behaviour) are counted on the device and neutralised (index 0, weight 0).  ``check_indices=True``
    reports them WITHOUT a host sync: the counter travels to pinned host memory behind an event and the
    next layer call / backward that finds the event finished raises IndexError (``BaryPack.validate()``
    waits for it); ``check_indices="sync"`` raises here, at the price of one host sync per new tensor;
    ``False`` skips the report.
    """
    _need_cuda(bary, "barycentric coordinates")
    if bary.dim() != 5 or bary.shape[3] != 3 or bary.shape[4] != 2:
        raise ValueError(f"barycentric coordinates must have shape (V,R,A,3,2), got {tuple(bary.shape)}")
    if bary.dtype not in (torch.float32, torch.float64):
        bary = bary.to(torch.float32)
    v, r, a = bary.shape[:3]
    idx, w, oob = prep_bary_op(bary, v_src)
    pack = BaryPack(idx=idx, w=w, v_out=v, v_src=v_src, r=r, a=a, keeps_all_slots=keep_zero_weight_slots)
    _PACK_OF_IDX[idx.data_ptr()] = pack
    if idx.numel() and check_indices:
        INDEX_CHECKS.submit(pack, oob)
        if check_indices == "sync":
            pack.validate()
    return pack


def prep_bary_traced(bary: torch.Tensor, v_src: int) -> BaryPack:
    """The split inside a torch.compile trace: one operator call, no cache, no host-side report."""
    if bary.dtype not in (torch.float32, torch.float64):
        bary = bary.to(torch.float32)
    idx, w, _oob = prep_bary_op(bary, v_src)
    return BaryPack(idx=idx, w=w, v_out=bary.shape[0], v_src=v_src, r=bary.shape[1], a=bary.shape[2])


class BaryCache:
    """Per-layer-instance-independent cache: bary tensor object -> BaryPack.

    Keyed on the tensor *object* (weak reference) and its version counter, never on data_ptr
    alone: the caching allocator recycles addresses."""

    def __init__(self, capacity: int = 16):
        self.capacity = capacity
        self._items: dict[int, tuple] = {}

    def get(self, bary: torch.Tensor, v_src: int, check_indices: bool = True) -> BaryPack:
        attached = getattr(bary, "_gcb_pack", None)      # a loader that packed the tensor itself (attach_pack)
        if attached is not None and attached[1] == bary._version and attached[0].v_src == v_src:
            return attached[0]
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


def attach_pack(bary: torch.Tensor, pack: BaryPack) -> torch.Tensor:
    """Tell the layers that `bary` has already been split / inverted: they use `pack` instead of running
    gcb_prep_bary and gcb_csr_build again (valid until `bary` is modified in place).  A data loader that keeps the
    meshes of a dataset on the device packs each of them once (data/faust_data_set.py::DeviceFaustCache)."""
    bary._gcb_pack = (pack, bary._version)
    return bary


def pack_with_weights(base: BaryPack, w: torch.Tensor) -> BaryPack:
    """A pack that shares `base`'s indices and CSR inversion but carries other weights (training-time noise on the
    barycentric weights: preprocess / faust_data_set.py:82-94 perturb the weight column only).  `base` must have been
    inverted WITHOUT dropping zero-weight slots (`prep_bary(..., keep_zero_weight_slots=True)`), because a weight that
    is zero in one epoch need not be in the next."""
    if not base.keeps_all_slots:
        raise ValueError("pack_with_weights needs a base pack built with keep_zero_weight_slots=True")
    out = BaryPack(idx=base.idx, w=w.contiguous(), v_out=base.v_out, v_src=base.v_src, r=base.r, a=base.a,
                   oob_count=base.oob_count, keeps_all_slots=True)
    out._csr = base.csr()
    _PACK_OF_IDX[out.idx.data_ptr()] = out
    return out


# ================================================================================================
# torch.library operators (namespace ``geoconv_b200``)
#
# Every compute entry point of the C ABI is registered as a custom operator with a fake (meta) kernel
# and an autograd formula, so a model that uses the drop-in layers traces under torch.compile with
# ``fullgraph=True`` (the reference demo compiles its model, training_demo.py:126-127) and the eager
# path and the compiled path run the same code.  The operator bodies only move pointers across the ABI.
# ================================================================================================
_LIB = "geoconv_b200"


def _direct(op, probe=None):
    """The Python body behind a custom operator.  Outside a torch.compile trace (and off fake tensors) the layers call
    the bodies directly (under a plain autograd.Function where a gradient is needed): the dispatcher adds 20-40 us of
    host time per operator call, which a step that is read back every iteration pays in full
    (profiles/r02_notes.md, section 15)."""
    if torch.compiler.is_compiling() or (probe is not None and _is_fake(probe)):
        return op
    return op._init_fn


_NO_GUARD = contextlib.nullcontext()


def _dev_guard(t: torch.Tensor):
    """Device guard for the C-ABI calls; nothing to do (and nothing to pay) when t's device is already current."""
    idx = t.device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(idx)


@torch.library.custom_op(f"{_LIB}::prep_bary", mutates_args=(), device_types="cuda")
def prep_bary_op(bary: torch.Tensor, v_src: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(V,R,A,3,2) fp32/fp64 -> idx int32 (V,R,A,3), w fp32 (V,R,A,3), out-of-range counter int32 (1,)."""
    lib = N.load()
    bary = bary.contiguous()
    v, r, a = bary.shape[:3]
    idx = torch.empty((v, r, a, 3), dtype=torch.int32, device=bary.device)
    w = torch.empty((v, r, a, 3), dtype=torch.float32, device=bary.device)
    oob = torch.zeros(1, dtype=torch.int32, device=bary.device)
    if idx.numel():
        with _dev_guard(bary):
            N.check(lib.gcb_prep_bary(bary.data_ptr(), int(bary.dtype == torch.float64), idx.numel(), v_src,
                                      idx.data_ptr(), w.data_ptr(), oob.data_ptr(), _stream(bary)),
                    "gcb_prep_bary")
    return idx, w, oob


@prep_bary_op.register_fake
def _(bary, v_src):
    v, r, a = bary.shape[:3]
    return (bary.new_empty((v, r, a, 3), dtype=torch.int32), bary.new_empty((v, r, a, 3), dtype=torch.float32),
            bary.new_empty((1,), dtype=torch.int32))


@torch.library.custom_op(f"{_LIB}::csr_build", mutates_args=(), device_types="cuda")
def csr_build_op(idx: torch.Tensor, w: torch.Tensor, v_src: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Neighbour -> vertex inversion: (row_ptr int32 [V_src+1], entries int32 [n_slots], stats int32 [8]:
    [1] max |w| as fp32 bits, [2] longest row - what the fp16-split backward bounds its operand with)."""
    lib = N.load()
    n_slots = idx.numel()
    dev = idx.device
    r, a = idx.shape[1], idx.shape[2]
    row_ptr = torch.empty(v_src + 1, dtype=torch.int32, device=dev)
    entries = torch.empty(n_slots, dtype=torch.int32, device=dev)
    stats = torch.empty(8, dtype=torch.int32, device=dev)
    ws_bytes = lib.gcb_csr_workspace_bytes(n_slots, v_src)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with _dev_guard(idx):
        # (template size given: rows grouped by radial ring - _native.CSR_RING_GROUPED goes to the backward)
        N.check(lib.gcb_csr_build(idx.data_ptr(), w.data_ptr(), n_slots, v_src, r, a, row_ptr.data_ptr(),
                                  entries.data_ptr(), stats.data_ptr(), ws.data_ptr(), ws_bytes, _stream(idx)),
                "gcb_csr_build")
    return row_ptr, entries, stats


@csr_build_op.register_fake
def _(idx, w, v_src):
    return (idx.new_empty((v_src + 1,), dtype=torch.int32), idx.new_empty((idx.numel(),), dtype=torch.int32),
            idx.new_empty((8,), dtype=torch.int32))


@torch.library.custom_op(f"{_LIB}::pack_records", mutates_args=(), device_types="cuda")
def pack_records_op(idx: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """(idx, w) -> the slot-major, tile-padded gather records of the tensor-core forward (gcb_fwd_pack_records)."""
    lib = N.load()
    v_out, r, a = idx.shape[:3]
    out = torch.empty(lib.gcb_fwd_records_bytes(v_out, r, a), dtype=torch.uint8, device=idx.device)
    with _dev_guard(idx):
        N.check(lib.gcb_fwd_pack_records(idx.data_ptr(), w.data_ptr(), v_out, r, a, out.data_ptr(), _stream(idx)),
                "gcb_fwd_pack_records")
    return out


@pack_records_op.register_fake
def _(idx, w):
    v_out, r, a = idx.shape[:3]
    v_pad = (v_out + 127) // 128 * 128
    return idx.new_empty(((r * a * v_pad * 32 + 255) // 256 * 256 + 256,), dtype=torch.uint8)


# weight-ring schedules + K-split boundaries per (device, shape, rotation list, precision): built once
_FWD_TABLES: dict[tuple, torch.Tensor] = {}


def forward_tables(device, v_out, r, a, f, t, rot, precision):
    key = (device.index, v_out, r, a, f, t, tuple(rot), precision)
    tab = _FWD_TABLES.get(key)
    if tab is None:
        lib = N.load()
        tab = torch.empty(lib.gcb_fwd_tables_bytes(), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            N.check(lib.gcb_fwd_build_tables(N.rot_array(rot), len(rot), v_out, r, a, f, t, precision, tab.data_ptr(),
                                             torch.cuda.current_stream(device).cuda_stream), "gcb_fwd_build_tables")
        if len(_FWD_TABLES) > 64:
            _FWD_TABLES.clear()
        if not torch.cuda.is_current_stream_capturing():   # (a table built inside a capture belongs to that graph)
            _FWD_TABLES[key] = tab
    return tab


# ---- prior folding  W[T,R,A,F] x K[R,A,R,A] -> Weff -------------------------------------------
@torch.library.custom_op(f"{_LIB}::fold_prior", mutates_args=(), device_types="cuda")
def fold_prior_op(inp: torch.Tensor, prior: torch.Tensor, transpose: int) -> torch.Tensor:
    lib = N.load()
    inp = inp.contiguous()
    prior = prior.contiguous()
    t, r, a, f = inp.shape
    out = torch.empty_like(inp)
    with _dev_guard(inp):
        N.check(lib.gcb_fold_prior(inp.data_ptr(), prior.data_ptr(), out.data_ptr(), t, r, a, f, transpose,
                                   _stream(inp)), "gcb_fold_prior")
    return out


@fold_prior_op.register_fake
def _(inp, prior, transpose):
    return torch.empty_like(inp, memory_format=torch.contiguous_format)


def _fold_setup(ctx, inputs, output):
    _inp, prior, transpose = inputs
    ctx.save_for_backward(prior)
    ctx.transpose = transpose


def _fold_backward(ctx, grad):
    (prior,) = ctx.saved_tensors
    return _direct(fold_prior_op, grad)(grad, prior, 1 - ctx.transpose), None, None


fold_prior_op.register_autograd(_fold_backward, setup_context=_fold_setup)


class FoldPrior:
    """``FoldPrior.apply(W, K)`` -> Weff (differentiable in W)."""

    @staticmethod
    def apply(w_neighbor: torch.Tensor, prior: torch.Tensor) -> torch.Tensor:
        _need_cuda(w_neighbor, "template weights")
        prior = prior.to(device=w_neighbor.device, dtype=torch.float32)
        if torch.compiler.is_compiling() or _is_fake(w_neighbor):
            return fold_prior_op(w_neighbor, prior, 0)
        return _EagerFold.apply(w_neighbor, prior, 0)


class _EagerFold(torch.autograd.Function):
    """fold_prior_op's body and autograd formula without the operator dispatcher (eager mode)."""

    @staticmethod
    def forward(ctx, inp, prior, transpose):
        out = fold_prior_op._init_fn(inp, prior, transpose)
        _fold_setup(ctx, (inp, prior, transpose), out)
        return out

    @staticmethod
    def backward(ctx, grad):
        return _fold_backward(ctx, grad)


# ---- the convolution (+ fused pooling outputs) --------------------------------------------------
@torch.library.custom_op(f"{_LIB}::conv_fwd", mutates_args=(), device_types="cuda")
def conv_fwd_op(signal: torch.Tensor, w_eff: torch.Tensor | None, w_self: torch.Tensor, bias: torch.Tensor,
                idx: torch.Tensor, w: torch.Tensor, w_anchor: torch.Tensor | None, records: torch.Tensor | None,
                rot: list[int], act_id: int, prec_fwd: int, prec_bwd: int,
                row0: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(signal, Weff|None, Wself, bias) -> (Y [V,n_rot,T], Z [V,T], argmax [V], stats int32 [8]): ConvIntrinsic.forward
    and the AngularMaxPooling of its result in one pass (include/geoconv_b200.h: gcb_conv_fwd).  `w_anchor`: the
    neighbour weights of an all-zero-prior layer (they take no part, but receive an all-zeros gradient);
    `records`: cached gather records of the barycentric tensor; `stats`: operand maxima for the backward."""
    lib = N.load()
    x = signal.contiguous()
    v_src, f = x.shape
    v_out, r, a = idx.shape[:3]
    t = w_self.shape[0]
    n_rot = len(rot)
    dev = x.device
    w_self_c = w_self.reshape(t, f).contiguous()
    bias_c = bias.reshape(t).contiguous()
    w_eff_c = None if w_eff is None else w_eff.contiguous()
    y = torch.empty((v_out, n_rot, t), dtype=torch.float32, device=dev)
    z = torch.empty((v_out, t), dtype=torch.float32, device=dev)
    arg = torch.empty((v_out,), dtype=torch.int32, device=dev)
    stats = torch.empty((8,), dtype=torch.int32, device=dev)
    ws_bytes = lib.gcb_conv_fwd_workspace_bytes(v_out, v_src, r, a, f, t, n_rot, prec_fwd)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    tensor_path = prec_fwd != N.PREC_FP32 and w_eff_c is not None and v_out > 0
    tables = forward_tables(dev, v_out, r, a, f, t, rot, prec_fwd) if tensor_path else None
    with _dev_guard(x):
        N.check(lib.gcb_conv_fwd(x.data_ptr(), idx.data_ptr(), w.data_ptr(), _ptr(w_eff_c), w_self_c.data_ptr(),
                                 bias_c.data_ptr(), N.rot_array(rot), n_rot, v_out, v_src, r, a, f, t, act_id,
                                 prec_fwd, y.data_ptr(), z.data_ptr(), arg.data_ptr(), None,
                                 _ptr(records) if tensor_path else None, _ptr(tables), stats.data_ptr(), row0,
                                 ws.data_ptr(), ws_bytes, _stream(x)), "gcb_conv_fwd")
    return y, z, arg, stats


@conv_fwd_op.register_fake
def _(signal, w_eff, w_self, bias, idx, w, w_anchor, records, rot, act_id, prec_fwd, prec_bwd, row0):
    v_out, t = idx.shape[0], w_self.shape[0]
    return (signal.new_empty((v_out, len(rot), t), dtype=torch.float32),
            signal.new_empty((v_out, t), dtype=torch.float32), signal.new_empty((v_out,), dtype=torch.int32),
            signal.new_empty((8,), dtype=torch.int32))


@torch.library.custom_op(f"{_LIB}::conv_bwd", mutates_args=(), device_types="cuda")
def conv_bwd_op(grad_y: torch.Tensor | None, grad_z: torch.Tensor | None, x: torch.Tensor,
                w_eff: torch.Tensor | None, w_self: torch.Tensor, y: torch.Tensor, z: torch.Tensor,
                arg: torch.Tensor, idx: torch.Tensor, w: torch.Tensor, row_ptr: torch.Tensor | None,
                entries: torch.Tensor | None, csr_stats: torch.Tensor | None, fwd_stats: torch.Tensor | None,
                rot: list[int], act_id: int, precision: int, need_dx: bool, stages: int,
                ws_in: torch.Tensor | None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Backward of conv_fwd: a gradient arriving through Z takes the AMP-sparse route (one rotation per
    vertex), a gradient arriving through Y is back-propagated one rotation at a time.
    Returns (dSignal [V_src,F] or an empty tensor, dWeff [T,R,A,F], dWself [T,F], dbias [T], workspace).
    stages (AMP-sparse route only): 3 = everything; 1 = up to the weight gradients, the returned workspace then goes
    into a stages = 2 call (`ws_in`) that produces dSignal - a data-parallel caller starts the all-reduce of the
    weight gradients in between (see `GradSync`)."""
    lib = N.load()
    v_src, f = x.shape
    v_out, r, a = idx.shape[:3]
    t = w_self.shape[0]
    dev = x.device
    n_rot = len(rot)
    has_nb = w_eff is not None
    d_x = torch.empty_like(x) if need_dx and stages != 1 else None
    w_stage = stages != 2
    staged = None
    if stages == 1 and GRAD_SYNC is not None and GRAD_SYNC.wants_staging(t % 4 == 0):
        # data parallel, peer-memory all-reduce: the weight gradients are written straight into this rank's
        # peer-readable staging area (no copy on either side of the collective)
        staged = GRAD_SYNC.peer_reducer.staging([t * r * a * f, t * f, t], dev)
    if staged is not None:
        d_w_eff, d_w_self, d_bias = staged[0].view(t, r, a, f), staged[1].view(t, f), staged[2]
    else:
        d_w_eff = torch.empty((t, r, a, f) if w_stage else (0,), dtype=torch.float32, device=dev)
        d_w_self = torch.empty((t, f) if w_stage else (0,), dtype=torch.float32, device=dev)
        d_bias = torch.empty((t,) if w_stage else (0,), dtype=torch.float32, device=dev)
    ws_bytes = lib.gcb_conv_amp_bwd_workspace_bytes(v_out, v_src, r, a, f, t, precision)
    ws = ws_in if ws_in is not None else torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = _stream(x)
    rot_c = N.rot_array(rot)
    calls = 0
    if grad_y is not None:
        h = torch.empty((v_out, t), dtype=torch.float32, device=dev)
        rot_sel = torch.empty((v_out,), dtype=torch.int32, device=dev)

    def run(accumulate):
        N.check(lib.gcb_conv_bwd(x.data_ptr(), idx.data_ptr(), w.data_ptr(), _ptr(w_eff) if has_nb else None,
                                 w_self.data_ptr(), h.data_ptr(), rot_sel.data_ptr(), _ptr(row_ptr), _ptr(entries),
                                 N.CSR_RING_GROUPED, v_out, v_src, r, a, f, t, precision, accumulate, None, _ptr(d_x),
                                 d_w_eff.data_ptr(), d_w_self.data_ptr(), d_bias.data_ptr(), ws.data_ptr(),
                                 ws_bytes, st), "gcb_conv_bwd")

    with _dev_guard(x):
        if grad_z is not None:
            # AMP-sparse route: one rotation per vertex, one call (activation gradient, rotation select, bias
            # reduction, operand scales and splits ride in its single preparation kernel)
            gz = grad_z.contiguous().to(torch.float32)
            N.check(lib.gcb_conv_amp_bwd(gz.data_ptr(), z.data_ptr(), arg.data_ptr(), rot_c, n_rot, act_id, x.data_ptr(),
                                         idx.data_ptr(), w.data_ptr(), _ptr(w_eff) if has_nb else None,
                                         w_self.data_ptr(), _ptr(row_ptr), _ptr(entries), N.CSR_RING_GROUPED,
                                         _ptr(csr_stats), _ptr(fwd_stats), v_out, v_src, r, a, f, t, precision, stages,
                                         _ptr(d_x), d_w_eff.data_ptr() if w_stage else 256,
                                         d_w_self.data_ptr() if w_stage else 256, d_bias.data_ptr() if w_stage else 256,
                                         ws.data_ptr(), ws_bytes, st), "gcb_conv_amp_bwd")
            calls += 1
        if grad_y is not None:
            gy = grad_y.contiguous().to(torch.float32)
            offs = [int(o) % a for o in rot]
            if (has_nb and row_ptr is not None and len(set(offs)) == n_rot and os.environ.get("GCB_DENSE_BWD", "1") != "0"
                    and lib.gcb_conv_bwd_dense_supported(r, a, f, t, precision)):
                # every rotation in ONE pass: h for the whole (V, n_rot, T) tensor, one G, one pair of GEMMs
                yc = y.contiguous()
                h_all = torch.empty_like(gy)
                N.check(lib.gcb_act_grad(gy.data_ptr(), yc.data_ptr(), gy.numel(), act_id, h_all.data_ptr(), st),
                        "gcb_act_grad")
                N.check(lib.gcb_conv_bwd_dense(x.data_ptr(), idx.data_ptr(), w.data_ptr(), w_eff.data_ptr(),
                                               w_self.data_ptr(), h_all.data_ptr(), N.rot_array(offs), n_rot,
                                               row_ptr.data_ptr(), entries.data_ptr(), N.CSR_RING_GROUPED, v_out, v_src, r, a,
                                               f, t, precision, 1 if calls else 0, _ptr(d_x), d_w_eff.data_ptr(),
                                               d_w_self.data_ptr(), d_bias.data_ptr(), ws.data_ptr(), ws_bytes, st),
                        "gcb_conv_bwd_dense")
                calls += 1
            else:
                for i, off in enumerate(rot):   # shapes the dense kernel does not tile: one sparse backward per rotation
                    gi = gy[:, i, :].contiguous()
                    yi = y[:, i, :].contiguous()
                    N.check(lib.gcb_act_grad(gi.data_ptr(), yi.data_ptr(), gi.numel(), act_id, h.data_ptr(), st),
                            "gcb_act_grad")
                    rot_sel.fill_(off % a)
                    run(1 if calls else 0)
                    calls += 1
    if d_x is None:
        d_x = x.new_empty((0,))
    if staged is not None:
        # (an operator's outputs may not share storage: the gradients stay where they were written - the caller
        # takes the views from the reducer - and empty placeholders are returned)
        d_w_eff, d_w_self, d_bias = x.new_empty((0,)), x.new_empty((0,)), x.new_empty((0,))
    return d_x, d_w_eff, d_w_self, d_bias, (ws if stages == 1 else x.new_empty((0,), dtype=torch.uint8))


@conv_bwd_op.register_fake
def _(grad_y, grad_z, x, w_eff, w_self, y, z, arg, idx, w, row_ptr, entries, csr_stats, fwd_stats, rot, act_id,
      precision, need_dx, stages, ws_in):
    v_out, r, a = idx.shape[:3]
    t, f = w_self.shape[0], x.shape[1]
    return (torch.empty_like(x) if need_dx else x.new_empty((0,)), x.new_empty((t, r, a, f)), x.new_empty((t, f)),
            x.new_empty((t,)), x.new_empty((0,), dtype=torch.uint8))


def _is_fake(t: torch.Tensor) -> bool:
    from torch._subclasses.fake_tensor import is_fake
    return is_fake(t)


class GradSync:
    """Data-parallel weight-gradient reduction started INSIDE the layer's backward (geoconv_b200/distributed.py installs
    one): the backward runs in two stages; after the first (weight gradients final) `reduce_early` launches their
    all-reduce asynchronously - the collective's stream waits for what has been enqueued so far, i.e. for stage one -
    and the dSignal GEMM of stage two runs beside it; `join` makes the compute stream wait for the collective."""

    def __init__(self, group=None, average: bool = False, peer_reducer=None):
        self.group = group
        self.average = average
        self.active = True
        self._works = []
        self.early_reductions = 0
        self.peer_reducer = peer_reducer      # distributed.PeerAllReduce: the library's own NVLink all-reduce kernel
        self.peer_failed = None
        self._side = None
        self._joined = True

    def wants_staging(self, sizes_ok: bool) -> bool:
        return bool(self.active and self.peer_reducer is not None and sizes_ok)

    def reduce_early(self, tensors):
        """Start the sum of `tensors` over the ranks; returns None (reduced in place once `join` has run) or, when
        the tensors are views of the peer reducer's staging area, a callable that returns private copies of the sums
        (call it after `join`)."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return None
        tensors = [t_ for t_ in tensors if t_.numel()]
        if self.peer_reducer is not None and tensors and tensors[0].is_cuda and len(tensors) <= 3 and \
                all(t_.numel() % 4 == 0 and t_.is_contiguous() for t_ in tensors):
            try:
                dev = tensors[0].device
                if self._side is None:
                    self._side = torch.cuda.Stream(device=dev, priority=-1)
                self._side.wait_stream(torch.cuda.current_stream(dev))      # stage one (weight gradients) is enqueued
                sizes = [t_.numel() for t_ in tensors]
                n = sum(sizes)
                stage = self.peer_reducer.staging(sizes, dev)
                if all(t_.data_ptr() == s_.data_ptr() for t_, s_ in zip(tensors, stage)):
                    self.peer_reducer.reduce_staged(n, dev, stream=self._side)
                    shapes = [tuple(t_.shape) for t_ in tensors]

                    def take():
                        flat = self.peer_reducer.result(n, dev).clone()
                        out, off = [], 0
                        for k, shp in zip(sizes, shapes):
                            out.append(flat[off:off + k].view(shp))
                            off += k
                        return out
                    self._joined = False
                    self.early_reductions += 1
                    return take
                self.peer_reducer(tensors, stream=self._side)
                self._joined = False
                self.early_reductions += 1
                return None
            except Exception as exc:  # noqa: BLE001 - e.g. symmetric memory unavailable: NCCL from now on
                self.peer_failed = f"{type(exc).__name__}: {exc}"
                self.peer_reducer = None
        done = False
        if len(tensors) > 1 and tensors[0].is_cuda:
            try:     # one NCCL launch for all of them (coalesced group)
                with dist._coalescing_manager(group=self.group, device=tensors[0].device, async_ops=True) as cm:
                    for t_ in tensors:
                        dist.all_reduce(t_, op=dist.ReduceOp.SUM, group=self.group)
                self._works.append(cm)
                done = True
            except Exception:  # noqa: BLE001 - private API: fall back to one collective per tensor
                done = False
        if not done:
            for t_ in tensors:
                self._works.append(dist.all_reduce(t_, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.early_reductions += 1
        return None

    def join(self):
        for w_ in self._works:
            w_.wait()          # (stream-level wait: the host does not block)
        self._works = []
        if not self._joined:
            torch.cuda.current_stream(self._side.device).wait_stream(self._side)
            self._joined = True


GRAD_SYNC: GradSync | None = None


# idx tensor -> BaryPack that owns it (eager mode: the CSR inversion is built once per barycentric tensor)
_PACK_OF_IDX: "weakref.WeakValueDictionary[int, BaryPack]" = weakref.WeakValueDictionary()


def _csr_for(idx: torch.Tensor, w: torch.Tensor, v_src: int):
    from torch._subclasses.fake_tensor import is_fake
    if not is_fake(idx):
        pack = _PACK_OF_IDX.get(idx.data_ptr())
        if pack is not None and pack.v_src == v_src and pack.idx.data_ptr() == idx.data_ptr():
            return pack.csr()
    return tuple(csr_build_op(idx, w, v_src))


def _conv_setup(ctx, inputs, output):
    signal, w_eff, w_self, bias, idx, w, w_anchor, _records, rot, act_id, prec_fwd, prec_bwd, row0 = inputs
    ctx.row0 = row0
    y, z, arg, stats = output
    # the forward's operand maxima serve the backward when both run the fp16-split kernels on the same operands
    ctx.use_fwd_stats = prec_fwd == N.PREC_F16X3 and prec_bwd == N.PREC_F16X3 and w_eff is not None
    ctx.rot = [int(o) for o in rot]
    ctx.act_id = act_id
    ctx.precision = prec_bwd
    ctx.has_neighbor = w_eff is not None
    ctx.has_anchor = w_anchor is not None
    ctx.anchor_shape = None if w_anchor is None else tuple(w_anchor.shape)
    ctx.w_self_shape = tuple(w_self.shape)
    ctx.bias_shape = tuple(bias.shape)
    t, f = w_self.shape[0], signal.shape[1]
    ctx.save_for_backward(signal, w_eff if w_eff is not None else signal.new_empty((0,)),
                          w_self.reshape(t, f), y, z, arg, idx, w, stats)
    ctx.set_materialize_grads(False)


def _conv_backward(ctx, grad_y, grad_z, _grad_arg, _grad_stats):
    x, w_eff, w_self, y, z, arg, idx, w, stats = ctx.saved_tensors
    if grad_y is None and grad_z is None:
        return (None,) * 13
    if ctx.row0:
        raise NotImplementedError("geoconv_b200: a convolution over a row block (row_offset > 0) is forward-only")
    INDEX_CHECKS.poll()
    has_nb = ctx.has_neighbor
    need_dx = bool(ctx.needs_input_grad[0])
    # the CSR inversion also feeds the weight gradients (dW = G^T X on tensor cores), not only dSignal
    row_ptr, entries, csr_stats = _csr_for(idx, w, x.shape[0]) if has_nb else (None, None, None)
    args = (grad_y, grad_z, x.contiguous(), w_eff if has_nb else None, w_self, y, z, arg, idx, w, row_ptr, entries,
            csr_stats, stats if ctx.use_fwd_stats else None, ctx.rot, ctx.act_id, ctx.precision, need_dx)
    sync = GRAD_SYNC
    if (sync is not None and sync.active and grad_y is None and need_dx and has_nb and not _is_fake(x) and
            N.load().gcb_conv_amp_bwd_can_split(idx.shape[1], idx.shape[2], x.shape[1], w_self.shape[0], ctx.precision, 1)):
        # data parallel: weight gradients first, their all-reduce (collective stream) beside the dSignal GEMM
        _none, d_w_eff, d_w_self, d_bias, ws = _direct(conv_bwd_op, x)(*args, 1, None)
        if d_w_eff.numel() == 0:     # written into the peer reducer's staging area (see conv_bwd_op)
            t_, r_, a_, f_ = w_self.shape[0], idx.shape[1], idx.shape[2], x.shape[1]
            st_ = sync.peer_reducer.staging([t_ * r_ * a_ * f_, t_ * f_, t_], x.device)
            d_w_eff, d_w_self, d_bias = st_[0].view(t_, r_, a_, f_), st_[1].view(t_, f_), st_[2]
        staged = sync.reduce_early([d_w_eff, d_w_self, d_bias])
        d_x = _direct(conv_bwd_op, x)(*args, 2, ws)[0]
        sync.join()
        if staged is not None:       # the sums sit in the collective's result area: private copies for autograd
            d_w_eff, d_w_self, d_bias = staged()
    else:
        d_x, d_w_eff, d_w_self, d_bias, _ws = _direct(conv_bwd_op, x)(*args, 3, None)
    g_anchor = torch.zeros(ctx.anchor_shape, dtype=torch.float32, device=x.device) if ctx.has_anchor else None
    return (d_x if need_dx else None, d_w_eff if has_nb else None, d_w_self.reshape(ctx.w_self_shape),
            d_bias.reshape(ctx.bias_shape), None, None, g_anchor, None, None, None, None, None, None)


conv_fwd_op.register_autograd(_conv_backward, setup_context=_conv_setup)


class ConvIntrinsicFn:
    """``apply(signal, Weff|None, Wself, bias, pack, rot_offsets, act_id, precision, w_anchor=None)``
    -> (Y, Z, argmax).  `precision` is one C-ABI code for both directions or a (forward, backward) pair."""

    @staticmethod
    def apply(signal, w_eff, w_self, bias, pack: BaryPack, rot_offsets, act_id, precision, w_anchor=None, row0=0):
        _need_cuda(signal, "signal")
        prec_fwd, prec_bwd = precision if isinstance(precision, tuple) else (precision, precision)
        records = None
        if prec_fwd != N.PREC_FP32 and w_eff is not None and not torch.compiler.is_compiling() and pack.v_out > 0:
            records = pack.records()
        args = (signal, w_eff, w_self, bias, pack.idx, pack.w, w_anchor, records, [int(o) for o in rot_offsets], act_id,
                prec_fwd, prec_bwd, int(row0))
        traced = torch.compiler.is_compiling() or _is_fake(signal)
        y, z, arg, _stats = conv_fwd_op(*args) if traced else _EagerConv.apply(*args)
        return y, z, arg


class _EagerConv(torch.autograd.Function):
    """conv_fwd_op's body and autograd formula without the operator dispatcher (eager mode)."""

    @staticmethod
    def forward(ctx, *inputs):
        out = conv_fwd_op._init_fn(*inputs)
        _conv_setup(ctx, inputs, out)
        ctx.mark_non_differentiable(out[2], out[3])
        return out

    @staticmethod
    def backward(ctx, grad_y, grad_z, _grad_arg, _grad_stats):
        return _conv_backward(ctx, grad_y, grad_z, _grad_arg, _grad_stats)


# ---- angular poolings on a materialised tensor ---------------------------------------------------
POOL_MAX, POOL_MIN, POOL_AVG = 0, 1, 2


@torch.library.custom_op(f"{_LIB}::apool_fwd", mutates_args=(), device_types="cuda")
def apool_fwd_op(y: torch.Tensor, mode: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Angular max / min / average pooling of (V, n_rot, T): (Z (V,T), selected rotation int32 (V,))."""
    lib = N.load()
    yc = y.contiguous().to(torch.float32)
    v, n_rot, t = yc.shape
    z = torch.empty((v, t), dtype=torch.float32, device=yc.device)
    arg = torch.empty((v,), dtype=torch.int32, device=yc.device)
    with _dev_guard(yc):
        if mode == POOL_MAX:
            N.check(lib.gcb_amp_fwd(yc.data_ptr(), v, n_rot, t, z.data_ptr(), arg.data_ptr(), _stream(yc)), "gcb_amp_fwd")
        else:
            N.check(lib.gcb_apool_fwd(yc.data_ptr(), v, n_rot, t, mode, z.data_ptr(), arg.data_ptr(), _stream(yc)),
                    "gcb_apool_fwd")
    return z, arg


@apool_fwd_op.register_fake
def _(y, mode):
    v, _n_rot, t = y.shape
    return y.new_empty((v, t), dtype=torch.float32), y.new_empty((v,), dtype=torch.int32)


@torch.library.custom_op(f"{_LIB}::apool_bwd", mutates_args=(), device_types="cuda")
def apool_bwd_op(grad_z: torch.Tensor, arg: torch.Tensor, n_rot: int, mode: int) -> torch.Tensor:
    lib = N.load()
    gz = grad_z.contiguous().to(torch.float32)
    v, t = gz.shape
    gy = torch.empty((v, n_rot, t), dtype=torch.float32, device=gz.device)
    with _dev_guard(gz):
        if mode == POOL_MAX:
            N.check(lib.gcb_amp_bwd(gz.data_ptr(), arg.data_ptr(), v, n_rot, t, gy.data_ptr(), _stream(gz)), "gcb_amp_bwd")
        else:
            N.check(lib.gcb_apool_bwd(gz.data_ptr(), arg.data_ptr(), v, n_rot, t, mode, gy.data_ptr(), _stream(gz)),
                    "gcb_apool_bwd")
    return gy


@apool_bwd_op.register_fake
def _(grad_z, arg, n_rot, mode):
    v, t = grad_z.shape
    return grad_z.new_empty((v, n_rot, t), dtype=torch.float32)


def _apool_setup(ctx, inputs, output):
    y, mode = inputs
    ctx.save_for_backward(output[1])
    ctx.n_rot = y.shape[1]
    ctx.mode = mode


def _apool_backward(ctx, grad_z, _grad_arg):
    (arg,) = ctx.saved_tensors
    return apool_bwd_op(grad_z, arg, ctx.n_rot, ctx.mode), None


apool_fwd_op.register_autograd(_apool_backward, setup_context=_apool_setup)


class AngularPoolFn:
    """``apply(y, mode)`` -> (Z, selected rotation); mode = POOL_MAX / POOL_MIN / POOL_AVG."""

    @staticmethod
    def apply(y, mode):
        _need_cuda(y, "inputs")
        return apool_fwd_op(y, mode)


class AngularMaxPoolFn:
    @staticmethod
    def apply(y):
        _need_cuda(y, "inputs")
        return apool_fwd_op(y, POOL_MAX)


@torch.library.custom_op(f"{_LIB}::pack_rows", mutates_args=(), device_types="cuda")
def pack_rows_op(src: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    lib = N.load()
    src = src.contiguous()
    rows = rows.to(device=src.device, dtype=torch.int32).contiguous()
    out = torch.empty((rows.numel(), src.shape[1]), dtype=torch.float32, device=src.device)
    with _dev_guard(src):
        N.check(lib.gcb_pack_rows(src.data_ptr(), rows.data_ptr(), rows.numel(), src.shape[1], out.data_ptr(),
                                  _stream(src)), "gcb_pack_rows")
    return out


@pack_rows_op.register_fake
def _(src, rows):
    return src.new_empty((rows.numel(), src.shape[1]), dtype=torch.float32)


def pack_rows(src: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """out[i] = src[rows[i]] (halo send-buffer packing)."""
    _need_cuda(src, "src")
    return pack_rows_op(src, rows)


# ---- between two convolutions: BatchNorm1d -> Dropout --------------------------
_BN_WS: dict = {}


def _bn_workspace(v: int, t: int, dev: torch.device) -> torch.Tensor:
    """Column-reduction scratch; zeroed once (the kernels leave the ticket zero), grown on demand, per device.
    One scratch per device: calls on the same device are expected on ONE stream at a time (like the layers'
    workspaces, which come from the caching allocator of the current stream)."""
    need = N.load().gcb_bn_workspace_bytes(v, t)
    ws = _BN_WS.get(dev)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=dev)
        if not torch.cuda.is_current_stream_capturing():
            _BN_WS[dev] = ws
    return ws


@torch.library.custom_op(f"{_LIB}::bn_stats", mutates_args=("running_mean", "running_var"), device_types="cuda")
def bn_stats_op(z: torch.Tensor, eps: float, running_mean: torch.Tensor | None, running_var: torch.Tensor | None,
                momentum: float) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-feature (mean, 1/sqrt(biased var + eps)) of z (V, T), fixed summation order; the running estimates, when
    given, are updated in place by the same kernel (torch's rule, unbiased variance)."""
    lib = N.load()
    z = z.contiguous()
    v, t = z.shape
    mean, invstd = (torch.empty(t, dtype=torch.float32, device=z.device) for _ in range(2))
    ws = _bn_workspace(v, t, z.device)
    with _dev_guard(z):
        N.check(lib.gcb_bn_stats(z.data_ptr(), v, t, float(eps), mean.data_ptr(), invstd.data_ptr(),
                                 _ptr(running_mean), _ptr(running_var), float(momentum), ws.data_ptr(), ws.numel(),
                                 _stream(z)), "gcb_bn_stats")
    return mean, invstd


@bn_stats_op.register_fake
def _(z, eps, running_mean, running_var, momentum):
    t = z.shape[1]
    return z.new_empty((t,)), z.new_empty((t,))


@torch.library.custom_op(f"{_LIB}::bn_dropout", mutates_args=(), device_types="cuda")
def bn_dropout_op(z: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                  training: bool, p: float, seed: int, stream_id: int) -> torch.Tensor:
    """dropout_p(gamma * (z - mean) * invstd + beta) in one pass; the mask is a function of (seed, stream_id, element)."""
    lib = N.load()
    z = z.contiguous()
    v, t = z.shape
    out = torch.empty_like(z)
    with _dev_guard(z):
        N.check(lib.gcb_bn_dropout_fwd(z.data_ptr(), v, t, mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                       beta.data_ptr(), float(p), seed, stream_id, out.data_ptr(), _stream(z)),
                "gcb_bn_dropout_fwd")
    return out


@bn_dropout_op.register_fake
def _(z, mean, invstd, gamma, beta, training, p, seed, stream_id):
    return torch.empty_like(z)


@torch.library.custom_op(f"{_LIB}::bn_dropout_bwd", mutates_args=(), device_types="cuda")
def bn_dropout_bwd_op(d_out: torch.Tensor, z: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor, gamma: torch.Tensor,
                      training: bool, p: float, seed: int, stream_id: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    lib = N.load()
    d_out, z = d_out.contiguous(), z.contiguous()
    v, t = z.shape
    d_z = torch.empty_like(z)
    d_gamma, d_beta = (torch.empty(t, dtype=torch.float32, device=z.device) for _ in range(2))
    ws = _bn_workspace(v, t, z.device)
    with _dev_guard(z):
        N.check(lib.gcb_bn_dropout_bwd(d_out.data_ptr(), z.data_ptr(), v, t, mean.data_ptr(), invstd.data_ptr(),
                                       gamma.data_ptr(), int(training), float(p), seed, stream_id, d_z.data_ptr(),
                                       d_gamma.data_ptr(), d_beta.data_ptr(), ws.data_ptr(), ws.numel(), _stream(z)),
                "gcb_bn_dropout_bwd")
    return d_z, d_gamma, d_beta


@bn_dropout_bwd_op.register_fake
def _(d_out, z, mean, invstd, gamma, training, p, seed, stream_id):
    t = z.shape[1]
    return torch.empty_like(z), z.new_empty((t,)), z.new_empty((t,))


def _bn_dropout_setup(ctx, inputs, output):
    z, mean, invstd, gamma, _beta, training, p, seed, stream_id = inputs
    ctx.save_for_backward(z, mean, invstd, gamma)
    ctx.cfg = (bool(training), float(p), int(seed), int(stream_id))


def _bn_dropout_backward(ctx, d_out):
    z, mean, invstd, gamma = ctx.saved_tensors
    training, p, seed, stream_id = ctx.cfg
    # in training mode mean / invstd are functions of z: the kernel's d_z is the full derivative, so they get none
    d_z, d_gamma, d_beta = bn_dropout_bwd_op(d_out, z, mean, invstd, gamma, training, p, seed, stream_id)
    return d_z, None, None, d_gamma, d_beta, None, None, None, None


bn_dropout_op.register_autograd(_bn_dropout_backward, setup_context=_bn_dropout_setup)
