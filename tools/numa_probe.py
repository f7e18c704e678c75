"""H2D bandwidth of one GPU against the NUMA node the host buffer lives on, pinned (cudaHostAlloc) vs registered in place.
   python tools/numa_probe.py <gpu> ; prints GB/s for 2 MB and 64 MB copies from every NUMA node."""
import os, sys, glob, time
import numpy as np
import torch

gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.cuda.set_device(gpu)
props = torch.cuda.get_device_properties(gpu)
bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
try:
    node_of_gpu = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
except Exception as e:
    node_of_gpu = f"? ({e})"
nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
print(f"gpu {gpu} pci {bus} numa_node {node_of_gpu}; nodes {nodes}", flush=True)
cudart = torch.cuda.cudart()

def cpus_of(node):
    s = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in s.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out

dst = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for node in nodes:
    cpus = cpus_of(node)
    os.sched_setaffinity(0, cpus)
    for kind in ("pinned", "registered"):
        for size in (2 << 20, 64 << 20):
            if kind == "pinned":
                h = torch.empty(size, dtype=torch.uint8).pin_memory()
                h.fill_(1)
            else:
                a = np.ones(size, dtype=np.uint8)
                h = torch.from_numpy(a)
                assert cudart.cudaHostRegister(a.ctypes.data, size, 1) == 0 or True
            torch.cuda.synchronize()
            reps = 40 if size < (8 << 20) else 10
            for _ in range(3): dst[:size].copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps): dst[:size].copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"  node {node} ({len(cpus)} cpus) {kind:10s} {size >> 20:3d} MB: {size * reps / dt * 1e-9:6.1f} GB/s", flush=True)
            if kind == "registered":
                cudart.cudaHostUnregister(a.ctypes.data)
