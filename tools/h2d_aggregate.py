"""Aggregate host->device bandwidth of a box: one process per GPU, all copying pinned host memory at the same time.
   python tools/h2d_aggregate.py [ngpus] [MB per copy]"""
import os, sys, time, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    gpu, mb, t_start = int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
    torch.cuda.set_device(gpu)
    size = mb << 20
    h = torch.empty(size, dtype=torch.uint8).pin_memory(); h.fill_(1)
    d = torch.empty(size, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    while time.time() < t_start: pass
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 2.0:
        for _ in range(8): d.copy_(h, non_blocking=True)
        torch.cuda.synchronize(); n += 8
    dt = time.perf_counter() - t0
    print(f"gpu {gpu}: {size * n / dt * 1e-9:.1f} GB/s", flush=True)
else:
    import torch
    ng = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    mb = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    for group in ([0], list(range(ng))):
        t_start = time.time() + 25
        ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "child", str(g), str(mb), str(t_start)], stdout=subprocess.PIPE, text=True) for g in group]
        outs = [p.communicate()[0].strip() for p in ps]
        tot = sum(float(o.split(":")[1].split()[0]) for o in outs)
        print(f"{len(group)} GPU(s) at once, {mb} MB copies: " + "; ".join(outs) + f"  => total {tot:.1f} GB/s", flush=True)
