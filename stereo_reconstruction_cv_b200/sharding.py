"""Frame sharding of a batch of stereo pairs over the GPUs of one box.

The dense-stereo path shards by stereo pair / video frame only: vertical and diagonal SGM paths
span the whole image, so there is nothing to exchange between GPUs while a frame is computed.
One process per GPU (torchrun), a contiguous block of the batch per rank, NO collective on the
data path.  The only optional exchange is the gather of the compacted point clouds (or the
disparity maps) to one rank at the end: counts first, then a variable-length gather.

The helpers take any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU
tests (tests/test_sharding.py, world_size 2).
"""
import numpy as np


def shard_range(n_frames, world_size, rank):
    """Contiguous block partition frames[g*B/G : (g+1)*B/G] of SURVEY.md 8(e) -> (start, stop)."""
    if world_size < 1 or not (0 <= rank < world_size) or n_frames < 0:
        raise ValueError("bad shard request: n_frames=%r world_size=%r rank=%r" % (n_frames, world_size, rank))
    return (rank * n_frames) // world_size, ((rank + 1) * n_frames) // world_size


def shard_ranges(n_frames, world_size):
    return [shard_range(n_frames, world_size, r) for r in range(world_size)]


def compute_shard(stereo, lefts, rights, world_size=1, rank=0, out=None):
    """Disparity of this rank's block of a batch.

    lefts / rights: sequences (or (B,H,W) arrays / CUDA tensors) indexed by global frame number.
    Returns (start, stop, list_or_tensor_of_disparities).  CUDA tensors stay on the device."""
    start, stop = shard_range(len(lefts), world_size, rank)
    if hasattr(lefts, "is_cuda") and lefts.is_cuda and lefts.dim() >= 3:
        if start == stop:                                 # more ranks than frames: an empty block, not an error
            import torch
            return start, stop, torch.empty((0,) + tuple(lefts.shape[1:3]), dtype=torch.int16, device=lefts.device)
        return start, stop, stereo.compute(lefts[start:stop], rights[start:stop], out)
    if start == stop:
        return start, stop, []
    # host arrays: ONE call of the batched host entry point for the whole block, so that the H2D / D2H
    # copies of neighbouring frames overlap the kernels and small frames run side by side
    if hasattr(stereo, "compute_batch"):
        lb = np.stack([np.asarray(lefts[i]) for i in range(start, stop)])
        rb = np.stack([np.asarray(rights[i]) for i in range(start, stop)])
        res = stereo.compute_batch(lb, rb, out)
        return start, stop, [res[i] for i in range(stop - start)]
    return start, stop, [stereo.compute(lefts[i], rights[i]) for i in range(start, stop)]


def compute_batch_devices(params, lefts, rights, devices=None, out=None):
    """One process, several GPUs: the batch is block-partitioned over `devices` (default: every visible CUDA device) and
    every block goes through ONE `compute_batch` call on its own device, from its own host thread (SURVEY.md 8(e):
    "one worker -- process or thread -- per GPU").  No collective, no peer traffic: each thread owns a StereoSGBM handle
    bound to its device (sgbm_create binds a handle to the current device; the library serialises per handle and keeps
    per-device kernel set-up under a lock, so threads on different devices do not interact), and ctypes releases the
    GIL for the duration of the call.

    params: dict of StereoSGBM_create arguments.  lefts / rights: uint8 numpy (B,H,W) or (B,H,W,3), ideally page-locked.
    Returns the int16 (B,H,W) disparities in frame order (written into `out` when given)."""
    import threading

    import torch

    from .stereo import StereoSGBM_create
    lefts = np.ascontiguousarray(lefts)
    rights = np.ascontiguousarray(rights)
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    if not devices:
        raise RuntimeError("compute_batch_devices needs at least one CUDA device: the engine has no CPU fallback")
    B, H, W = lefts.shape[:3]
    if out is None:
        out = np.empty((B, H, W), np.int16)
    errors = [None] * len(devices)

    def worker(k, dev):
        try:
            start, stop = shard_range(B, len(devices), k)
            if start == stop:
                return
            with torch.cuda.device(dev):
                st = StereoSGBM_create(**params)          # bound to `dev`
                st.compute_batch(lefts[start:stop], rights[start:stop], out[start:stop])
                del st
        except Exception as e:                            # surfaced in the calling thread
            errors[k] = e

    threads = [threading.Thread(target=worker, args=(k, d)) for k, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    return out


def gather_varlen(t, dst=0, group=None):
    """Gather first-dimension-ragged tensors (e.g. per-rank point clouds, N_r x 3) to rank `dst`.

    Two steps: all-gather of the per-rank lengths (8 bytes each), then a gather
    of buffers padded to the maximum length.  Returns the list of per-rank tensors on `dst` and
    None elsewhere.  Without an initialised process group it degenerates to [t]."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [t]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts) if counts else 0
    pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return [b[:c] for b, c in zip(bufs, counts)]


def gather_point_cloud(xyz, rgb=None, dst=0, group=None):
    """Concatenate the ranks' compacted clouds (reprojectCompact output) on rank `dst` in rank order,
    i.e. in global frame order for block-sharded batches.  Returns (xyz, rgb) on dst, (None, None) elsewhere."""
    import torch
    px = gather_varlen(xyz, dst, group)
    pc = gather_varlen(rgb, dst, group) if rgb is not None else None
    if px is None:
        return None, None
    return torch.cat(px, 0), (torch.cat(pc, 0) if pc is not None else None)
