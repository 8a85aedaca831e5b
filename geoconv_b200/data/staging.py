"""Host-side staging for the inputs of a step: ONE pinned, write-combined arena, ONE host-to-device copy.

Why.  Measured on the B200 boxes (tools/h2d_probe2.py, profiles/r02_h2d_probe2.jsonl): a 24 MB pinned buffer that the
CPU has just written is copied at 22-34 GB/s - its lines sit dirty in the CPU caches and every DMA read has to snoop
them - while write-combined pinned memory (``cudaHostAllocWriteCombined``: never cached, not snooped) is copied at
43-48 GB/s and an untouched buffer at 54 GB/s.  And every separate copy pays a fixed cost, so the signal, the
barycentric tensor and the upstream gradient of a step travel as one piece.

``PinnedArena([t0, t1, ...])`` allocates one arena that holds tensors shaped like ``t_i`` (each 256-byte aligned) and
exposes them as ``.views`` (CPU tensors: WRITE into them, do not read - reads of write-combined memory are slow);
``.buffer`` is the whole arena as a uint8 tensor.  ``device_arena`` builds the matching device-side arena, so a step's
inputs arrive with ``dev.buffer.copy_(host.buffer, non_blocking=True)``.
"""
import ctypes
import glob
import os

import torch

_WRITE_COMBINED = 0x04   # cudaHostAllocWriteCombined
_rt = None


def _cudart():
    global _rt
    if _rt is None:
        cands = sorted(glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")))
        cands += sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"))
        if not cands:
            raise RuntimeError("geoconv_b200: libcudart not found")
        _rt = ctypes.CDLL(cands[0])
    return _rt


def _layout(tensors):
    sizes = [(t.numel() * t.element_size() + 255) // 256 * 256 for t in tensors]
    return sizes, sum(sizes)


def _views(buffer, tensors, sizes):
    out, off = [], 0
    for t, n in zip(tensors, sizes):
        out.append(buffer[off:off + t.numel() * t.element_size()].view(t.dtype).view(t.shape))
        off += n
    return out


class PinnedArena:
    def __init__(self, like, write_combined=True):
        torch.cuda.init()
        self._sizes, self.nbytes = _layout(like)
        self._ptr = ctypes.c_void_p()
        rc = _cudart().cudaHostAlloc(ctypes.byref(self._ptr), ctypes.c_size_t(max(self.nbytes, 256)),
                                     ctypes.c_uint(_WRITE_COMBINED if write_combined else 0))
        if rc != 0:
            raise RuntimeError(f"geoconv_b200: cudaHostAlloc failed with error {rc}")
        self._raw = (ctypes.c_uint8 * max(self.nbytes, 256)).from_address(self._ptr.value)
        self.buffer = torch.frombuffer(self._raw, dtype=torch.uint8)[:self.nbytes]
        self.views = _views(self.buffer, like, self._sizes)
        self.write_combined = bool(write_combined)

    def fill(self, tensors):
        """Writes one step's inputs into the arena (sequential stores: what write-combined memory is good at)."""
        for dst, src in zip(self.views, tensors):
            dst.copy_(src)
        return self

    def __del__(self):
        ptr, self._ptr = getattr(self, "_ptr", None), None
        if ptr is not None and ptr.value:
            try:
                _cudart().cudaFreeHost(ptr)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass


def device_arena(like, device):
    """(buffer uint8 on ``device``, views shaped like ``like``) with the layout of ``PinnedArena(like)``."""
    sizes, total = _layout(like)
    buffer = torch.empty(total, dtype=torch.uint8, device=device)
    return buffer, _views(buffer, like, sizes)
