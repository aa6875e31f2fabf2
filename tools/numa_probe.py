"""Concurrent device -> host read-back bandwidth of every GPU of the box, with the host buffer allocated (a) wherever the
process happens to run and (b) after the process is pinned to the CPUs of the GPU's own NUMA node (first touch there).
torchrun --nproc-per-node N tools/numa_probe.py"""
import os, time
import torch, torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def gpu_numa_node(ordinal):
    bdf = torch.cuda.get_device_properties(ordinal)
    try:
        import ctypes
        rt = torch.cuda.cudart()
        # PCI bus id through the runtime: domain:bus:device.function
        buf = ctypes.create_string_buffer(32)
        lib = ctypes.CDLL("libcudart.so", mode=ctypes.RTLD_GLOBAL) if False else None
    except Exception:
        pass
    ident = f"{bdf.pci_domain_id:04x}:{bdf.pci_bus_id:02x}:{bdf.pci_device_id:02x}.0"
    try:
        node = int(open(f"/sys/bus/pci/devices/{ident}/numa_node").read())
    except OSError:
        node = -1
    return ident, node


def node_cpus(node):
    try:
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    except OSError:
        return None
    cpus = []
    for part in txt.split(","):
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    return cpus


def bandwidth(nbytes=256 << 20, reps=8):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.zero_()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    return nbytes * reps / dt / 1e9


ident, node = gpu_numa_node(local)
before = bandwidth()
cpus = node_cpus(node) if node >= 0 else None
allowed = sorted(os.sched_getaffinity(0))
if cpus:
    keep = [c for c in cpus if c in allowed]
    if keep:
        os.sched_setaffinity(0, keep)
after = bandwidth()
line = f"rank {rank} gpu {ident} numa {node} cpus_allowed {len(allowed)} local_cpus {len(cpus) if cpus else None} d2h GB/s default {before:.1f} pinned-to-node {after:.1f}"
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, line)
    if rank == 0:
        print("\n".join(out))
        os.system("nvidia-smi topo -m 2>/dev/null | head -14; ls /sys/devices/system/node/ | head; nproc")
else:
    print(line)
