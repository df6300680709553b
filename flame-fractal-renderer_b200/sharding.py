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


def slice_bounds(cells, cell, world):
    """Element range [first, first + n) of the buffer slice each rank reduces: whole cells,
    ceil(cells / world) per rank (the rule ffr_cuda_reduce uses between devices)."""
    per = (cells + world - 1) // world
    out = []
    for r in range(world):
        c0 = min(per * r, cells)
        c1 = min(c0 + per, cells)
        out.append((c0 * cell, (c1 - c0) * cell))
    return out


def _typed_sum_torch(dst, srcs, first_elem, cell):
    """dst += sum(srcs), element j of a cell typed by position (0: u64 count, else f64 colour
    sum) -- the torch form of K2d, for host tensors only (the gloo tests); device tensors go
    through the library's kernel."""
    n = dst.numel()
    pos = (torch.arange(n, dtype=torch.int64) + first_elem) % cell
    is_count = pos == 0
    acc_i = dst.clone()
    acc_f = dst.view(torch.float64).clone()
    for s in srcs:
        acc_i += s
        acc_f += s.view(torch.float64)
    dst.copy_(torch.where(is_count, acc_i, acc_f.view(torch.int64)))


def reduce_buffer(buf, cells, cell, dst=0, group=None, renderer=None):
    """Sum every rank's raw buffer (int64 view of the reference layout, cells x cell elements)
    into rank `dst`. Integer counts are exact and order independent; colour sums are f64.

    cell == 1 (counts only): one `reduce` of the int64 view (u64 sum == i64 sum mod 2^64).
    cell > 1: the layout interleaves u64 counts with f64 colour sums, which no single typed
    collective can add, so the exchange is a reduce-scatter built from an all-to-all of buffer
    slices + the library's typed slice sum (K2d reduce_slices_kernel through
    ffr_cuda_sum_device_slices, `renderer` = the rank's BufferRenderer) + sends of the finished
    slices to `dst`: no de-interleaving copies, every rank moves (N-1)/N of a buffer each way.
    Works on NCCL (GPU tensors, needs `renderer`) and on gloo (host tensors, CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return buf
    assert buf.dtype == torch.int64 and buf.numel() == cells * cell
    if cell == 1:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=group)
        return buf
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bounds = slice_bounds(cells, cell, world)
    first, n = bounds[rank]
    # slice `rank` of every peer arrives here (own slice included: all_to_all sends to self too)
    recv = torch.empty(n * world, dtype=torch.int64, device=buf.device)
    dist.all_to_all_single(recv, buf, output_split_sizes=[n] * world,
                           input_split_sizes=[b[1] for b in bounds], group=group)
    mine = buf[first:first + n]
    peers = [recv[j * n:(j + 1) * n] for j in range(world) if j != rank]
    if n:
        if buf.is_cuda:
            if renderer is None:
                raise RuntimeError("reduce_buffer: colour buffers on the GPU need the renderer "
                                   "(typed slice sum kernel); there is no torch fallback on device")
            if renderer.own_stream:
                # the library launches on a stream of its own: order it after the all-to-all
                torch.cuda.current_stream(buf.device).synchronize()
            renderer.sum_device_slices(mine.data_ptr(), [p.data_ptr() for p in peers], first, n)
            if renderer.own_stream:
                renderer.sync()
        else:
            _typed_sum_torch(mine, peers, first, cell)
    # finished slices travel to dst
    ops = []
    if rank == dst:
        for j in range(world):
            if j != dst and bounds[j][1]:
                ops.append(dist.P2POp(dist.irecv, buf[bounds[j][0]:bounds[j][0] + bounds[j][1]], j, group))
    elif n:
        ops.append(dist.P2POp(dist.isend, mine, dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return buf
