"""Does pinning the process to the GPU's NUMA-local cores (nvmlDeviceSetCpuAffinity) change the pinned H2D bandwidth?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml as nv

def bw(tag):
    x = torch.empty(160 * 1024 * 1024 // 8, dtype=torch.float64).pin_memory()
    x.fill_(1.0)
    d = torch.empty_like(x, device="cuda")
    for _ in range(2): d.copy_(x, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): d.copy_(x, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print(tag, f"{x.numel() * 8 / dt / 1e9:.1f} GB/s", "affinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)), "cpus", flush=True)

torch.cuda.init()
bw("default ")
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(0)
try:
    nv.nvmlDeviceSetCpuAffinity(h)
    bw("gpu-local")
except Exception as e:
    print("nvmlDeviceSetCpuAffinity failed:", e)
os.system("nvidia-smi topo -m 2>/dev/null | head -6; lscpu | grep -i 'numa\\|model name' | head -6")
