"""What limits the end-to-end (host buffers) figure at N GPUs?  Run under torchrun with N ranks.

Every rank, concurrently with all others (barrier before each leg):
  h2d   : 40 x cudaMemcpyAsync of one pinned 67 MB image to its GPU        -> GB/s per rank
  d2h   : 40 x copy of 10 MB of records into pinned host memory             -> GB/s per rank
  both  : the two together on two streams
  host  : 40 x the host-side work of one e2e step (numpy copy of 10 MB of records out of the pinned buffer)
Rank 0 prints one JSON line with min / mean / max over the ranks, plus the CPU affinity and NUMA layout it sees.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from sift_pyocl_b200 import _lib  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, REPS = 4096 * 4096, 40
himg = torch.from_numpy(_lib.pinned_empty((N,), np.float32))
himg.fill_(1.0)
hwc = torch.from_numpy(_lib.pinned_empty((N,), np.float32, True))  # write-combined
hwc.fill_(1.0)
dimg = torch.empty(N, dtype=torch.float32, device="cuda")
drec = torch.zeros(10_000_000, dtype=torch.uint8, device="cuda")
hrec = torch.from_numpy(_lib.pinned_empty((10_000_000,), np.uint8))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def leg(fn):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(REPS):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return dt / REPS


def h2d():
    with torch.cuda.stream(s1):
        dimg.copy_(himg, non_blocking=True)


def h2d_wc():
    with torch.cuda.stream(s1):
        dimg.copy_(hwc, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hrec.copy_(drec, non_blocking=True)


def both():
    h2d()
    d2h()


out_np = np.empty(10_000_000, np.uint8)
hrec_np = hrec.numpy()


def host():
    np.copyto(out_np, hrec_np)


res = {"h2d_GBps": N * 4 / leg(h2d) / 1e9, "h2d_wc_GBps": N * 4 / leg(h2d_wc) / 1e9, "d2h_GBps": 1e7 / leg(d2h) / 1e9, "both_ms": 1e3 * leg(both),
       "host_copy_ms": 1e3 * leg(host)}
keys = sorted(res)
t = torch.tensor([res[k] for k in keys], dtype=torch.float64, device="cuda")
if world > 1:
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    allt = torch.stack(allt).cpu().numpy()
else:
    allt = t.cpu().numpy()[None]
if rank == 0:
    line = {"n_gpus": world, "affinity_rank0": sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))],
            "cpu_count": os.cpu_count()}
    try:
        line["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    except OSError:
        line["numa_nodes"] = None
    for i, k in enumerate(keys):
        line[k] = {"min": float(allt[:, i].min()), "mean": float(allt[:, i].mean()), "max": float(allt[:, i].max())}
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()
