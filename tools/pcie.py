# raw H2D bandwidth of the box (pinned host memory, one large copy and 1 MiB chunks on 1/4 streams): the e2e ceiling
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for label, chunks, streams in [("one 1 GiB copy", 1, 1), ("1 MiB chunks, 1 stream", 1024, 1), ("1 MiB chunks, 4 streams", 1024, 4), ("8 MiB chunks, 4 streams", 128, 4)]:
    ss = [torch.cuda.Stream() for _ in range(streams)]
    best = 0
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        step = n // chunks
        for i in range(chunks):
            with torch.cuda.stream(ss[i % streams]):
                d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = max(best, n / dt / 1e9)
    print(f"{label}: {best:.1f} GB/s", flush=True)
