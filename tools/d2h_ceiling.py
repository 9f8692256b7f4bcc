#!/usr/bin/env python
"""What the box's PCIe / host memory can take: k GPUs copying device -> page-locked host memory at the same time (torchrun, one rank per GPU).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/d2h_ceiling.py

For k = 1, 2, 4, .. WORLD active ranks: every active rank copies `--mb` MiB per step (default: its share of a 66 MB compact 4K frame, and a
64 MiB block), 30 steps, into (a) its own cudaHostAlloc buffer and (b) its bands' share of ONE POSIX-shm buffer registered in every process
(what bench.py's e2e leg uses).  Prints per-rank and aggregate GB/s (wall clock between barriers, max over ranks).  The e2e figures of
bench.py at N GPUs are to be read against these ceilings."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
frame_bytes = 3840 * 2160 * 8


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(host, dev, active, steps=30):
    st = torch.cuda.current_stream()
    barrier()
    t0 = time.perf_counter()
    if rank < active:
        for _ in range(steps):
            host.copy_(dev, non_blocking=True)
        st.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt if rank < active else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / steps


from multiprocessing import shared_memory

names = [None]
shm_bytes = 64 << 20
if rank == 0:
    shm = shared_memory.SharedMemory(create=True, size=shm_bytes * world)
    names[0] = shm.name
if world > 1:
    dist.broadcast_object_list(names, src=0)
if rank != 0:
    shm = shared_memory.SharedMemory(name=names[0])
    try:
        from multiprocessing import resource_tracker

        resource_tracker.unregister(shm._name, "shared_memory")
    except Exception:
        pass
whole = np.ndarray((shm_bytes * world,), dtype=np.uint8, buffer=shm.buf)
mine = whole[rank * shm_bytes : (rank + 1) * shm_bytes]
mine[:] = 0  # first touch by the rank that will receive into it
barrier()
rc = int(torch.cuda.cudart().cudaHostRegister(whole.ctypes.data, shm_bytes * world, 0))
shm_t = torch.from_numpy(mine)
k = 1
while k <= world:
    for label, nbytes in (("frame share", frame_bytes // k), ("64 MiB", 64 << 20)):
        nbytes = min(nbytes, shm_bytes)
        dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        own = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        for _ in range(2):
            run(own, dev, k, 3)
        s_own = run(own, dev, k)
        for _ in range(2):
            run(shm_t[:nbytes], dev, k, 3)
        s_shm = run(shm_t[:nbytes], dev, k)
        if rank == 0:
            print(f"{k} GPU(s) x {nbytes / 1e6:7.2f} MB ({label}): own pinned buffer {nbytes / s_own / 1e9:6.1f} GB/s per GPU, {k * nbytes / s_own / 1e9:6.1f} aggregate ({s_own * 1e3:.3f} ms/step) | "
                  f"shared registered frame (cudaHostRegister rc {rc}) {nbytes / s_shm / 1e9:6.1f} per GPU, {k * nbytes / s_shm / 1e9:6.1f} aggregate ({s_shm * 1e3:.3f} ms/step)", flush=True)
        del dev, own
    k *= 2
barrier()
torch.cuda.cudart().cudaHostUnregister(whole.ctypes.data)
del shm_t, mine, whole
shm.close()
if rank == 0:
    shm.unlink()
if world > 1:
    dist.destroy_process_group()
