"""Host-to-device / device-to-host copy rate from pinned memory against the copy size (torch, one stream)."""
import torch, time
for mb in (1, 2, 4, 16, 64, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    reps = max(4, 512 // mb)
    for direction in ("h2d", "d2h"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("%s %4d MB x %3d: %.1f GB/s" % (direction, mb, reps, n * reps / dt * 1e-9), flush=True)
