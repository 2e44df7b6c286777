#!/usr/bin/env python
"""Host->device ceiling of the box for the end-to-end leg of bench.py: the product's own upload path
(mfk_upload_frames = cudaMemcpyAsync from page-locked memory on the context's copy stream) with NO kernels queued,
from the same kind of HostBuffer and in the same chunk sizes as bench.py's end-to-end leg, on N ranks concurrently.

  python profiles/h2d_ceiling.py                                   one rank
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/h2d_ceiling.py

Prints one JSON line per variant from rank 0: GB/s per GPU (min / mean over ranks; every rank times its own copies with
CUDA events between two barriers) for
  plain        the bench's chunking (64-frame chunks of 720x576 = 26.5 MB), default CPU affinity
  one_copy     the whole step (425 MB) as one copy
  affinity     the rank's thread pinned to the CPUs NVML names as local to its GPU (before the pinned allocation)
  torch        torch.Tensor.copy_(pinned, non_blocking=True) of the same bytes: an independent cross-check
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

W, H = 720, 576
FRAMES = 1024                # frames per step per GPU (bench.py: 16 streams x 64 new frames)
CHUNK = 64                   # frames per upload call
REPS = 12


def main():
    import torch
    import torch.distributed as dist
    import mofreak_b200

    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [float(o[0]) for o in out]
        return [x]

    frame = W * H
    results = []

    def measure(name, host_ptr, ctx, chunk, note=""):
        def one_pass():
            for f0 in range(0, FRAMES, chunk):
                ctx.upload_raw(host_ptr + f0 * frame, min(chunk, FRAMES - f0), f0)
        for _ in range(2):
            one_pass()
        ctx.sync(); barrier()
        t0 = time.perf_counter()
        for _ in range(REPS):
            one_pass()
        ctx.sync()
        wall = time.perf_counter() - t0
        barrier()
        gbs = REPS * FRAMES * frame / wall / 1e9
        per = gather(gbs)
        if rank == 0:
            results.append({"variant": name, "n_gpus": world, "gb_per_s_per_gpu_min": min(per), "gb_per_s_per_gpu_mean": sum(per) / len(per),
                            "gb_per_s_aggregate": sum(per), "bytes_per_pass": FRAMES * frame, "chunk_bytes": chunk * frame, "reps": REPS,
                            "note": note})

    ctx = mofreak_b200.Context(W, H, FRAMES, 8, device=local)
    host = mofreak_b200.HostBuffer(FRAMES * frame)
    host.array[:] = np.random.default_rng(rank).integers(0, 256, host.array.size, dtype=np.uint8)
    measure("plain", host.ptr, ctx, CHUNK, "mfk_upload_frames, 64-frame chunks, default affinity")
    measure("one_copy", host.ptr, ctx, FRAMES, "mfk_upload_frames, one 425 MB copy per pass")

    # independent cross-check: torch's own pinned allocator and copy engine path
    src = torch.empty(FRAMES * frame, dtype=torch.uint8, pin_memory=True)
    src.random_(0, 256)
    dst = torch.empty(FRAMES * frame, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(REPS):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    per = gather(REPS * FRAMES * frame / wall / 1e9)
    if rank == 0:
        results.append({"variant": "torch", "n_gpus": world, "gb_per_s_per_gpu_min": min(per), "gb_per_s_per_gpu_mean": sum(per) / len(per),
                        "gb_per_s_aggregate": sum(per), "note": "torch pinned tensor -> cuda tensor copy_, one 425 MB copy per pass"})
    del src, dst

    # NUMA/CPU affinity named by NVML for this GPU, set BEFORE allocating (first touch decides the pages' node)
    note = ""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        note = "pinned to %d CPUs named by NVML" % len(cpus)
    except Exception as e:      # no NVML binding: the variant still runs, the note says why it is the same as plain
        note = "NVML affinity unavailable (%s)" % type(e).__name__
    host2 = mofreak_b200.HostBuffer(FRAMES * frame)
    host2.array[:] = 7
    measure("affinity", host2.ptr, ctx, CHUNK, note)

    if rank == 0:
        for r in results:
            print(json.dumps(r), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
