"""8-GPU probe: concurrent device->host copy bandwidth per rank, with default pinned memory and after
binding the process to the GPU's NUMA-local CPUs (NVML affinity).  torchrun --nproc-per-node N scripts/probe_d2h.py"""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def numa_of_pages(t):
    return None
def measure(tag, nbytes=1 << 30, reps=8):
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst.fill_(1)                                   # touch
    for _ in range(2): dst.copy_(src, non_blocking=True)
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): dst.copy_(src, non_blocking=True)
    e.record(); barrier()
    gbs = nbytes * reps / (s.elapsed_time(e) * 1e-3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    allv = [torch.zeros_like(t) for _ in range(world)]
    if world > 1: dist.all_gather(allv, t)
    else: allv = [t]
    if rank == 0:
        v = [float(x) for x in allv]
        print(f"{tag:34s} per-rank GB/s: {' '.join(f'{x:5.1f}' for x in v)}  total {sum(v):.1f}", flush=True)
    del src, dst
if rank == 0:
    os.system("nvidia-smi topo -m 2>&1 | head -30; lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'; "
              "cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; grep -i allowed /proc/self/status")
    print("sched_getaffinity:", sorted(os.sched_getaffinity(0)), flush=True)
barrier()
measure("default pinned (concurrent)")
# one rank at a time
for r in range(world):
    if r == rank:
        src = torch.empty(1 << 30, dtype=torch.uint8, device=dev); dst = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
        dst.copy_(src); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(4): dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(); print(f"rank {r} alone: {4 * (1 << 30) / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
        del src, dst
    barrier()
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(local)
try:
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 4)
    cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
except Exception as ex:
    cpus = set(); print("rank", rank, "nvml affinity failed:", ex, flush=True)
allowed = os.sched_getaffinity(0)
inter = cpus & allowed
print(f"rank {rank}: nvml cpus {len(cpus)} (min {min(cpus) if cpus else None}, max {max(cpus) if cpus else None}), allowed {len(allowed)}, inter {len(inter)}", flush=True)
if inter:
    os.sched_setaffinity(0, inter)
barrier()
measure("after NUMA binding (concurrent)")
# half the ranks only (0..3 then 4..7)
for grp in ((0, 1, 2, 3), (4, 5, 6, 7), (0, 2, 4, 6), (0, 4)):
    if world < 8: break
    src = torch.empty(1 << 30, dtype=torch.uint8, device=dev); dst = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src); barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    if rank in grp:
        for _ in range(8): dst.copy_(src, non_blocking=True)
    e.record(); barrier()
    t = torch.tensor([8 * (1 << 30) / (s.elapsed_time(e) * 1e-3) / 1e9 if rank in grp else 0.0], dtype=torch.float64, device=dev)
    allv = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(allv, t)
    if rank == 0: print("group", grp, " ".join(f"{float(x):5.1f}" for x in allv), "total", sum(float(x) for x in allv), flush=True)
    del src, dst
if world > 1: dist.destroy_process_group()
