"""Device -> host copy ceiling of the box, per rank and in aggregate, for the end-to-end
Evaluate (the Jacobian values are 5.57 GB on BAL L).  Run under torchrun on N GPUs (or plain
python for one); every rank copies `--mb` megabytes from its GPU into host memory at the same
time, with four kinds of destination:

  hostalloc       cudaHostAlloc (torch pin_memory)
  register        cb200_host_alloc (anonymous mmap) + cb200_host_pin (cudaHostRegister): what
                  Evaluator::CreateJacobian hands out
  register_numa   the same, allocated and first-touched after pinning the process to the CPUs
                  of the GPU's NUMA node (/sys/bus/pci/devices/<gpu>/local_cpulist)
  pageable        plain numpy memory (the reference's std::unique_ptr<double[]>)

Rank 0 prints one JSON line: per-kind GB/s per rank (min over ranks) and in aggregate.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=860)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import ceres_b200  # noqa: E402,F401
from ceres_b200 import binding as B  # noqa: E402

n = args.mb * 1_000_000 // 8
src = torch.ones(n, dtype=torch.float64, device="cuda")


def gpu_cpulist():
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:
        return None


def run(dst_tensor):
    for _ in range(2):
        dst_tensor.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dst_tensor.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    return 8 * n / dt / 1e9


res = {}
res["hostalloc"] = run(torch.empty(n, dtype=torch.float64).pin_memory())
pa = B.PinnedArray(n)
pa.array[:] = 0.0
res["register"] = run(torch.from_numpy(pa.array))
pa.free()
cpus = gpu_cpulist()
if cpus:
    old = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, cpus & old or old)
        pa = B.PinnedArray(n)
        pa.array[:] = 0.0
        res["register_numa"] = run(torch.from_numpy(pa.array))
        pa.free()
    finally:
        os.sched_setaffinity(0, old)
res["pageable"] = run(torch.from_numpy(np.zeros(n)))

keys = sorted(res)
t = torch.tensor([res[k] for k in keys], dtype=torch.float64, device="cuda")
lo, total = t.clone(), t.clone()
if world > 1:
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
if rank == 0:
    print(json.dumps({"n_gpus": world, "mb_per_rank": args.mb,
                      "per_rank_GBps_min": dict(zip(keys, [round(float(v), 2) for v in lo.cpu()])),
                      "aggregate_GBps": dict(zip(keys, [round(float(v), 2) for v in total.cpu()])),
                      "numa_cpus_rank0": len(cpus) if cpus else None,
                      "host_cpus": os.cpu_count()}), flush=True)
if world > 1:
    dist.destroy_process_group()
