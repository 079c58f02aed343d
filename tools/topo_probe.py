"""Host topology probe for the multi-GPU e2e path: NUMA node of every GPU, cores, and aggregate pinned D2H bandwidth with all GPUs copying at once."""
import os, subprocess, sys, time, glob
import torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    try: print(os.path.basename(n), open(n + "/cpulist").read().strip())
    except Exception as e: print(n, e)
ng = torch.cuda.device_count()
for i in range(ng):
    bus = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
    q = subprocess.run(["nvidia-smi", "-i", str(i), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    dev = q.lower()[4:] if q.lower().startswith("0000") else q.lower()
    try: node = open("/sys/bus/pci/devices/%s/numa_node" % dev).read().strip()
    except Exception as e: node = "? (%s)" % e
    print("gpu", i, q, "numa", node)
n = 1 << 29
hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(ng)]
ds = [torch.empty(n, dtype=torch.uint8, device="cuda:%d" % i) for i in range(ng)]
for k in (1, 2, 4, ng):
    for i in range(k): torch.cuda.synchronize(i)
    t = time.perf_counter()
    for _ in range(4):
        for i in range(k): hs[i].copy_(ds[i], non_blocking=True)
    for i in range(k): torch.cuda.synchronize(i)
    print("D2H %d GPUs at once: %.1f GB/s aggregate" % (k, 4 * k * n / (time.perf_counter() - t) / 1e9))
