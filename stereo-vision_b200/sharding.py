"""Frame sharding across ranks (one process per GPU).

Stereo pairs are independent units (the reference builds a fresh Elas per frame,
stereothread.cpp:113), so the multi-GPU story is frame-sharded data parallelism with no data-path
collective: frame i goes to rank i mod world.  The only collective is one broadcast of the parameter
block (the 23-field Elas::parameters POD) from rank 0 at start-up, plus a MAX all-reduce of the
elapsed time for the throughput report.  Works with any torch.distributed backend (NCCL on the GPU
box, gloo in the CPU tests).
"""
import ctypes as C

import torch
import torch.distributed as dist


def shard_frames(n_frames, rank, world):
    """Indices of the frames rank `rank` processes: round-robin, frame i -> rank i mod world."""
    return list(range(rank, n_frames, world))


def broadcast_params(params, device, src=0):
    """Every rank ends up with rank `src`'s parameter block.  `params` is a ctypes Structure."""
    raw = bytearray(bytes(params))
    blob = torch.frombuffer(raw, dtype=torch.uint8).clone().to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
    return type(params).from_buffer_copy(blob.cpu().numpy().tobytes())


def max_over_ranks(values, device):
    """Element-wise MAX over ranks of a list of floats (device-side timings)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def sum_over_ranks(values, device):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.cpu()]
