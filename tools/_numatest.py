import os, sys, time, glob
sys.path.insert(0, "/root/repo")
import torch
dev = torch.device("cuda", 0)
p = torch.cuda.get_device_properties(0)
bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
try:
    print("gpu", bdf, "numa_node", open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
except Exception as ex:
    print("no sysfs numa", ex)
def cpus(node):
    txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in txt.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out
dbuf = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
print("affinity at start:", len(os.sched_getaffinity(0)), "cpus; running on cpu", os.sched_getcpu() if hasattr(os, "sched_getcpu") else "?")
for node in (0, 1, 0, 1):
    os.sched_setaffinity(0, cpus(node))
    time.sleep(0.05)
    h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    for n in (12 << 20, 64 << 20):
        dbuf[:n].copy_(h[:n], non_blocking=True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): dbuf[:n].copy_(h[:n], non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print(f"node {node}: H2D {n >> 20} MB: {dt * 1e3:.3f} ms  {n / dt / 1e9:.1f} GB/s")
    del h
