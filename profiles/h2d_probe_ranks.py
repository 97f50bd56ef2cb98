"""Pinned host->device copy rate per rank when N ranks copy AT THE SAME TIME (the ceiling of the e2e leg at N GPUs).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/h2d_probe_ranks.py
Each rank copies 185 MB H2D (+ 45 MB D2H on a second stream) per step, all ranks between barriers; three placements of the pinned
buffers: wherever the process happens to run, bound to the NUMA node of the rank's GPU (orbx.bind_to_gpu_numa), and write-combined."""
import ctypes
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from multi_orbslam3_b200 import orbx  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NB_IN, NB_OUT = 185 * 2 ** 20, 45 * 2 ** 20
d_in = torch.empty(NB_IN, dtype=torch.uint8, device="cuda"); d_out = torch.empty(NB_OUT, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
cudart = ctypes.CDLL("libcudart.so.12")


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def measure(h_in, h_out, n=10):
    def step():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    for _ in range(3):
        step()
    barrier(); t0 = time.perf_counter()
    for _ in range(n):
        step()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / n
    tt = torch.tensor([t], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())


def report(tag, t):
    if rank == 0:
        print("%-44s %6.2f ms per step per rank -> %6.1f GB/s H2D per rank, %7.1f GB/s aggregate (+ D2H %5.1f GB/s)" %
              (tag, t * 1e3, NB_IN / t / 1e9, world * NB_IN / t / 1e9, world * NB_OUT / t / 1e9), flush=True)


if rank == 0:
    print("ranks %d, cpus %d, numa nodes %s" % (world, os.cpu_count(), sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))), flush=True)
h_in = torch.empty(NB_IN, dtype=torch.uint8).pin_memory(); h_out = torch.empty(NB_OUT, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
report("pinned, unbound process", measure(h_in, h_out))
del h_in, h_out
info = orbx.bind_to_gpu_numa(local)
print("rank %d: gpu %d pci %s numa node %s cpus %s" % (rank, local, info.get("pci"), info.get("node"), info.get("cpus")), flush=True)
h_in = torch.empty(NB_IN, dtype=torch.uint8).pin_memory(); h_out = torch.empty(NB_OUT, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
report("pinned, process bound to the GPU's NUMA node", measure(h_in, h_out))
# write-combined host buffer for the input (cudaHostAllocWriteCombined = 4)
p = ctypes.c_void_p()
rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(NB_IN), ctypes.c_uint(4))
if rc == 0:
    import numpy as np
    arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(NB_IN,))
    arr[:] = 1
    h_wc = torch.from_numpy(arr)
    report("write-combined input, bound", measure(h_wc, h_out))
    del h_wc, arr
    cudart.cudaFreeHost(p)
if world > 1:
    dist.destroy_process_group()
