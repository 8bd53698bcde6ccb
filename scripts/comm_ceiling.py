"""What the fabric can do for the gradient exchange of BAL L (13,491,488 doubles = 108 MB),
measured with torch.distributed / NCCL and plain peer copies, to put the engine's own exchange
next to a ceiling.  Run under torchrun on N GPUs; rank 0 prints one JSON line.

  all_reduce     108 MB, sum, doubles           (what round 1 used: every rank sends everything)
  all_gather     N x (108 MB / N)               (exclusive ranges only: the bytes that must move)
  small_reduce   1 MB all-reduce                (cameras + boundary points + cost)
  p2p_copy       rank 0 -> its neighbour, 13.5 MB and 108 MB, cudaMemcpyPeer
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 13_491_488


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


full = torch.zeros(N, dtype=torch.float64, device="cuda")
part = torch.zeros(N // world, dtype=torch.float64, device="cuda")
gathered = torch.zeros(N // world * world, dtype=torch.float64, device="cuda")
small = torch.zeros(131072, dtype=torch.float64, device="cuda")
out = {"n_gpus": world, "doubles": N}
out["all_reduce_108MB_ms"] = timed(lambda: dist.all_reduce(full))
out["all_gather_ms"] = timed(lambda: dist.all_gather_into_tensor(gathered, part))
out["all_reduce_1MB_ms"] = timed(lambda: dist.all_reduce(small))
if rank == 0 and torch.cuda.device_count() > 1:
    peer = (local + 1) % torch.cuda.device_count()
    dst_small = torch.zeros(N // 8, dtype=torch.float64, device=f"cuda:{peer}")
    dst_full = torch.zeros(N, dtype=torch.float64, device=f"cuda:{peer}")

    def p2p(dst, src):
        for _ in range(5):
            dst.copy_(src)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            dst.copy_(src)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / 20 * 1e3
    out["p2p_copy_13MB_ms"] = p2p(dst_small, full[:N // 8])
    out["p2p_copy_108MB_ms"] = p2p(dst_full, full)
    out["p2p_GBps_108MB"] = 8 * N / out["p2p_copy_108MB_ms"] / 1e6
dist.barrier()
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
