"""Multi-GPU sharding of a frame batch: frames are independent, so ranks own contiguous frame
ranges (stereo L/R pairs stay on one rank), there is no collective on the data path, and the only exchange
is the final gather of per-frame point lists to rank 0.  Works with any torch.distributed backend (NCCL on
the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def frame_range(rank: int, world: int, total: int, group: int = 2):
    """[lo, hi) of the frames rank `rank` owns; ranges are multiples of `group` (2 = stereo pair) except the tail."""
    groups = (total + group - 1) // group
    per, extra = divmod(groups, world)
    glo = rank * per + min(rank, extra)
    ghi = glo + per + (1 if rank < extra else 0)
    return min(glo * group, total), min(ghi * group, total)


def gather_point_lists(local_lists, dst: int = 0):
    """local_lists: list of [n_i, 2] int32 arrays for this rank's frames (frame order).  Returns on `dst` the
    list for all frames in global frame order (ranks own consecutive ranges), elsewhere None."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local_lists)
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    counts = torch.tensor([len(a) for a in local_lists], dtype=torch.int64, device=dev)
    nframes = torch.tensor([len(local_lists)], dtype=torch.int64, device=dev)
    all_nframes = [torch.zeros_like(nframes) for _ in range(world)]
    dist.all_gather(all_nframes, nframes)
    all_nframes = [int(t.item()) for t in all_nframes]
    # counts per frame, then the flat point arrays (variable sizes: pad to the max and trim)
    maxf = max(all_nframes)
    cpad = torch.zeros(maxf, dtype=torch.int64, device=dev)
    cpad[:len(counts)] = counts
    all_counts = [torch.zeros_like(cpad) for _ in range(world)] if rank == dst else None
    dist.gather(cpad, all_counts, dst=dst)
    flat = np.concatenate([np.asarray(a, dtype=np.int32).reshape(-1, 2) for a in local_lists], axis=0) \
        if local_lists else np.zeros((0, 2), np.int32)
    npts = torch.tensor([len(flat)], dtype=torch.int64, device=dev)
    all_npts = [torch.zeros_like(npts) for _ in range(world)]
    dist.all_gather(all_npts, npts)
    maxp = max(int(t.item()) for t in all_npts)
    ppad = torch.zeros((maxp, 2), dtype=torch.int32, device=dev)
    ppad[:len(flat)] = torch.from_numpy(flat).to(dev)
    all_pts = [torch.zeros_like(ppad) for _ in range(world)] if rank == dst else None
    dist.gather(ppad, all_pts, dst=dst)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        c = all_counts[r][:all_nframes[r]].cpu().numpy()
        p = all_pts[r].cpu().numpy()
        off = 0
        for n in c:
            out.append(p[off:off + n].copy())
            off += int(n)
    return out


def gather_points_tensors(points, counts, dst: int = 0):
    """Tensor form of the gather, for shards whose point lists already sit compacted on the device: `points` [n, 2] int32 = the
    points of this rank's frames one after the other (frame order), `counts` [frames] int32.  One all_gather of the two sizes,
    then padded gathers.  Returns on `dst` (points of all frames in global frame order, counts of all frames), elsewhere
    (None, None).  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return points, counts
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = points.device
    sizes = torch.tensor([points.shape[0], counts.shape[0]], dtype=torch.int64, device=dev)
    allsz = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(allsz, sizes)
    maxp, maxf = max(int(z[0]) for z in allsz), max(int(z[1]) for z in allsz)
    ppad = torch.zeros((maxp, 2), dtype=torch.int32, device=dev)
    ppad[:points.shape[0]] = points
    cpad = torch.zeros((maxf,), dtype=torch.int32, device=dev)
    cpad[:counts.shape[0]] = counts
    gp = [torch.empty_like(ppad) for _ in range(world)] if rank == dst else None
    gc = [torch.empty_like(cpad) for _ in range(world)] if rank == dst else None
    dist.gather(ppad, gp, dst=dst)
    dist.gather(cpad, gc, dst=dst)
    if rank != dst:
        return None, None
    return (torch.cat([g[:int(z[0])] for g, z in zip(gp, allsz)]), torch.cat([g[:int(z[1])] for g, z in zip(gc, allsz)]))
