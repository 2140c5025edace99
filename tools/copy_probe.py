"""Copy-only probe for the e2e scaling question: every rank copies the bench's pinned host frames to its GPU
in 32-frame pieces, nothing else, barrier on both sides, max over ranks.  What the box's host memory / PCIe tree delivers to N
GPUs at once is the ceiling of e2e frames/s whatever the kernels do.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/copy_probe.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if os.environ.get("LGX_AFFINITY") == "1":
    # pin this rank to the CPUs NVML reports as local to its GPU before the page-locked buffer is allocated (first touch)
    import pynvml
    pynvml.nvmlInit()
    words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local), (os.cpu_count() + 63) // 64)
    cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
    if cpus:
        os.sched_setaffinity(0, cpus)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, B, CH = 2448, 2048, 256, 32
host = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
host.random_(0, 255)
dev = torch.empty((3, CH, H, W), dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(2)]


def one_pass():
    for c in range(B // CH):
        with torch.cuda.stream(streams[c % 2]):
            dev[c % 3].copy_(host[c * CH:(c + 1) * CH], non_blocking=True)
    torch.cuda.synchronize()


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


one_pass()
res = {}
for label, reps in (("h2d", 5),):
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        one_pass()
    sync()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    res[label] = float(dt.item())
if rank == 0:
    gb = B * H * W / 1e9
    print(json.dumps({"probe": "pinned H2D copy only, 32-frame pieces, 256 frames per GPU per pass", "n_gpus": world,
                      "s_per_pass": res["h2d"], "GB_per_s_per_gpu": gb / res["h2d"], "GB_per_s_total": world * gb / res["h2d"],
                      "frames_per_s_ceiling": world * B / res["h2d"], "affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."],
                      "cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
