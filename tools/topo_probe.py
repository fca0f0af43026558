"""Host topology of the bench box and the PCIe ceiling with all GPUs copying at once (not part of the product).
Usage: python tools/topo_probe.py [n_gpus]   -- prints core count / affinity / NUMA layout, then per-GPU pinned H2D GB/s with
n_gpus processes copying concurrently, once with the default CPU placement and once bound to the GPU's NUMA node."""
import glob
import multiprocessing as mp
import os
import subprocess
import sys
import time


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def gpu_numa_nodes():
    out = {}
    q = sh("nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader")
    for line in q.splitlines():
        idx, bus = [s.strip() for s in line.split(",")]
        bus = bus.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        p = f"/sys/bus/pci/devices/{bus}/numa_node"
        out[int(idx)] = int(open(p).read()) if os.path.exists(p) else -1
    return out


def node_cpus(node):
    p = f"/sys/devices/system/node/node{node}/cpulist"
    if not os.path.exists(p):
        return None
    cpus = []
    for part in open(p).read().strip().split(","):
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    return cpus


def worker(gpu, bind, barrier, q):
    if bind is not None:
        allowed = os.sched_getaffinity(0)
        want = [c for c in bind if c in allowed]
        if want:
            os.sched_setaffinity(0, want)
    import torch
    torch.cuda.set_device(gpu)
    h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty_like(h, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    q.put((gpu, 4 * h.numel() / (time.perf_counter() - t0) / 1e9))


def run(n, binds):
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n), ctx.Queue()
    ps = [ctx.Process(target=worker, args=(g, binds.get(g), barrier, q)) for g in range(n)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in range(n))
    [p.join() for p in ps]
    return [round(res[g], 1) for g in range(n)]


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], "...")
    print(sh("lscpu | grep -i -E 'model name|socket|numa|^cpu\\(s\\)'"))
    print("cgroup cpu.max", sh("cat /sys/fs/cgroup/cpu.max"))
    print("mem", sh("free -g | head -2"))
    print(sh("nvidia-smi topo -m"))
    nodes = gpu_numa_nodes()
    print("gpu numa nodes", nodes)
    print("nodes", [(os.path.basename(p), open(p + "/cpulist").read().strip()) for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))])
    print("H2D GB/s per GPU, default placement:", run(n, {}))
    binds = {g: node_cpus(nodes.get(g, -1)) for g in range(n) if nodes.get(g, -1) >= 0}
    print("H2D GB/s per GPU, bound to the GPU's node:", run(n, binds))
