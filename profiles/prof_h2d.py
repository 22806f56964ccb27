"""Host-to-device bandwidth of the GPU box for the end-to-end number: one pinned 200 MB buffer
(the C2 embedding) copied as a whole and in chunks on one / two streams."""
import time, torch
n = 200_000_000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
def run(chunks, streams):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    step = n // chunks
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c in range(chunks):
        with torch.cuda.stream(ss[c % streams]):
            d[c * step:(c + 1) * step].copy_(h[c * step:(c + 1) * step], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0
for chunks, streams in ((1, 1), (1, 1), (4, 1), (4, 2), (16, 4)):
    t = min(run(chunks, streams) for _ in range(3))
    print(f"chunks={chunks} streams={streams}: {t*1e3:.2f} ms  {n/t/1e9:.1f} GB/s")
hp = torch.empty(n, dtype=torch.uint8)  # pageable
t0 = time.perf_counter(); d.copy_(hp); torch.cuda.synchronize(); print(f"pageable: {(time.perf_counter()-t0)*1e3:.2f} ms")
