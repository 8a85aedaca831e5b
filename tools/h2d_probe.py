"""Pinned host -> device copy bandwidth on this box: one stream against several, whole buffers against chunks."""
import json, torch
dev = "cuda:0"
MB = 1 << 20
def bench(total_mb, n_streams, chunk_mb, iters=20):
    n = int(total_mb * MB)
    src = torch.empty(n, dtype=torch.uint8).pin_memory()
    dst = torch.empty(n, dtype=torch.uint8, device=dev)
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    chunk = int(chunk_mb * MB)
    pieces = [(o, min(chunk, n - o)) for o in range(0, n, chunk)]
    def go():
        for i, (o, l) in enumerate(pieces):
            with torch.cuda.stream(streams[i % n_streams]):
                dst[o:o + l].copy_(src[o:o + l], non_blocking=True)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(iters):
        go()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return round(n / ms / 1e6, 1)
for total in (24.25, 256):
    for ns, ch in ((1, total), (2, total / 2), (3, total / 3), (4, total / 4), (4, 1), (8, total / 8)):
        print(json.dumps({"total_MB": total, "streams": ns, "chunk_MB": round(ch, 2), "GB_per_s": bench(total, ns, ch)}), flush=True)
