"""Host-side multi-GPU plumbing for one-process-per-GPU launches (torchrun + NCCL).

The path shards with no data-path collective: the chain index space is
cut into contiguous ranges, one per rank; every rank renders into a private full-size buffer;
ONE exchange step sums the private buffers into rank 0 at the end. The buffer interleaves u64
counts with f64 colour sums per cell (reference layout, buffer_renderer.hpp:20-29), so a
single typed collective cannot reduce it: counts and colours travel as two planes.
"""

import torch
import torch.distributed as dist


def step_chain_range(step_index, rank, world, chains_per_step):
    """First chain of the disjoint range rank `rank` renders in step `step_index` (weak
    scaling: every rank renders chains_per_step chains per step)."""
    return (step_index * world + rank) * chains_per_step


def split_chains(total_chains, world):
    """Contiguous (first, count) per rank for a job of total_chains chains (strong split, the
    ffr-buf --gpus path and ffr_cuda_render_chains use the same rule)."""
    per = (total_chains + world - 1) // world
    out = []
    for r in range(world):
        first = min(per * r, total_chains)
        out.append((first, min(per, total_chains - first)))
    return out


def reduce_buffer(buf, cells, cell, dst=0, group=None):
    """Sum every rank's raw buffer (int64 view of the reference layout, cells x cell elements)
    into rank `dst`. Integer counts are exact and order independent; colour sums are f64.
    Works on any backend (NCCL on GPUs, gloo in the CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return buf
    assert buf.dtype == torch.int64 and buf.numel() == cells * cell
    if cell == 1:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=group)  # u64 sum == i64 sum mod 2^64
        return buf
    v = buf.view(cells, cell)
    counts = v[:, 0].contiguous()
    colors = v[:, 1:].contiguous().view(torch.float64)
    dist.reduce(counts, dst=dst, op=dist.ReduceOp.SUM, group=group)
    dist.reduce(colors, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if dist.get_rank(group) == dst:
        v[:, 0] = counts
        v[:, 1:] = colors.view(torch.int64)
    return buf
