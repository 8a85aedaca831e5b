"""One 24 MB pinned H2D copy per ~1 ms (the duty cycle of the e2e loop): default pinned memory against write-combined
pinned memory (cudaHostAllocWriteCombined), copy time measured per copy with events."""
import ctypes, glob, json, os, torch
dev = "cuda:0"
torch.cuda.init()
n = int(24.25 * (1 << 20))
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = ctypes.CDLL(cands[0])
def host_alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_uint8 * n).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8), p
dst = torch.empty(n, dtype=torch.uint8, device=dev)
copy = torch.cuda.Stream()
spin = int(0.55e-3 * 1.9e9)
for name, flags in (("default pinned (torch.pin_memory)", None), ("cudaHostAlloc default", 0), ("cudaHostAlloc write-combined", 4)):
    if flags is None:
        src = torch.empty(n, dtype=torch.uint8).pin_memory()
    else:
        src, keep = host_alloc(flags)
        src.fill_(1)
    times = []
    with torch.cuda.stream(copy):
        for i in range(60):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(copy)
            dst.copy_(src, non_blocking=True)
            e1.record(copy)
            torch.cuda._sleep(spin)          # the link idles for ~0.55 ms, like between two steps' copies
            times.append((e0, e1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in times]
    gbs = [round(n / t / 1e6, 1) for t in ms]
    print(json.dumps({"host memory": name, "GB_per_s copies 0-4": gbs[:5], "copies 25-29": gbs[25:30], "copies 55-59": gbs[55:]}), flush=True)
