"""Multi-GPU plumbing: one process per GPU, contiguous shards of the batch axis, no data-path collective.

Every frame / flow pair of the hot path is independent (no halo, no reduction: SURVEY section 8e), so the only
communication is (a) the max-over-ranks of a timing and (b) an OPTIONAL gather of the per-rank outputs to one rank over
NCCL / NVLink. Both go through torch.distributed (backend 'nccl' on GPUs, 'gloo' in the CPU tests); torch is imported
lazily so the single-GPU path does not depend on it.
"""
import numpy as np

from .batch import shard_range

__all__ = ['shard_range', 'rank_world', 'max_over_ranks', 'gather_frames', 'as_torch', 'bind_to_gpu_cpus']


def bind_to_gpu_cpus(device_index, local_rank=None, local_world=None):
    """Pin this process to CPU cores NVML reports as local to GPU `device_index` (same NUMA node / PCIe root), so that
    pinned host buffers allocated afterwards and the copy-issuing threads sit next to the GPU. With `local_rank` /
    `local_world` the GPU-local cores are dealt out round-robin, every rank of the node getting its own disjoint share
    (on a VM that reports the same core set for every GPU, 8 ranks bound to the same 32 cores contend for them: the
    end-to-end collapse of round 1). Returns the CPU list (empty if NVML or the affinity call is unavailable: then
    nothing is changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * wi + b for wi, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        cpus = [c for c in cpus if c < n_cpu]
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in cpus if c in allowed)
        if local_rank is not None and local_world and local_world > 1 and len(cpus) >= local_world:
            cpus = cpus[int(local_rank) % int(local_world)::int(local_world)]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []


def rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def as_torch(arr):
    """Zero-copy torch view of a DeviceArray (through __cuda_array_interface__) or of a numpy array."""
    import torch
    if isinstance(arr, np.ndarray):
        return torch.from_numpy(arr)
    return torch.as_tensor(arr, device='cuda')


def max_over_ranks(values):
    """Element-wise maximum of a small list of floats over all ranks (device timings are reported as the slowest
    rank's). Works on the process group's backend: CUDA tensors for nccl, CPU tensors for gloo."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def gather_frames(local, total_frames, dst=0):
    """Gather the per-rank shards [n_local, ...] of a batch of `total_frames` frames on rank `dst` (NCCL gather over
    NVLink for device arrays, gloo for numpy arrays). Shards are the contiguous ranges of shard_range(). Returns the
    assembled torch tensor on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    t = as_torch(local)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [shard_range(total_frames, r, world) for r in range(world)]
    assert t.shape[0] == sizes[rank][1] - sizes[rank][0], "local shard does not match shard_range()"
    out = None
    if rank == dst:
        out = torch.empty((total_frames,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    # point-to-point copies keep ragged shards simple; batched so NCCL runs them as one group over the NVSwitch
    ops = []
    if rank == dst:
        for r, (a, b) in enumerate(sizes):
            if r == dst:
                out[a:b].copy_(t)
            elif b > a:
                ops.append(dist.P2POp(dist.irecv, out[a:b], r))
    elif t.shape[0] > 0:
        ops.append(dist.P2POp(dist.isend, t.contiguous(), dst))
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    return out
