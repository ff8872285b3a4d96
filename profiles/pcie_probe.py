"""H2D bandwidth probe (GPU box): one pinned 708 MB buffer uploaded as 1 / 8 / 64 chunks on 1 or 2 streams."""
import time
import torch

n = 708_000_000 // 4
host = torch.empty(n, dtype=torch.float32).pin_memory()
dev = torch.empty(n, dtype=torch.float32, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]


def run(chunks, n_streams, reps=5):
    step = (n + chunks - 1) // chunks
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for c in range(chunks):
            with torch.cuda.stream(streams[c % n_streams]):
                dev[c * step:(c + 1) * step].copy_(host[c * step:(c + 1) * step], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"chunks {chunks:3d} streams {n_streams}: {best * 1e3:7.3f} ms  {n * 4 / best / 1e9:6.2f} GB/s", flush=True)


for chunks in (1, 8, 64):
    for ns in (1, 2, 4):
        if ns <= chunks:
            run(chunks, ns)
d2h = torch.empty(1920 * 1080 * 3, dtype=torch.float32).pin_memory()
src = torch.empty(1920 * 1080 * 3, dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    d2h.copy_(src, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print(f"D2H image: {dt * 1e3:.3f} ms {d2h.numel() * 4 / dt / 1e9:.2f} GB/s")
