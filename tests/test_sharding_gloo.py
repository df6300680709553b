"""The N > 1 host path on CPU: world_size-2 gloo. Two ranks take disjoint chain ranges
(sharding.step_chain_range / split_chains), fill private buffers, and the mixed u64/f64 buffer
reduce (sharding.reduce_buffer) must reproduce the single-process result: counts bit-exact,
colour sums to rounding. The per-rank renderer here is the oracle standing in for the device
(test infrastructure); the GPU tests check the device against the same oracle."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, size, chains, L, q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ffr = importlib.import_module("flame-fractal-renderer_b200")
    ex = importlib.import_module("flame-fractal-renderer_b200.examples")
    sharding = importlib.import_module("flame-fractal-renderer_b200.sharding")
    import pyoracle as po
    fl = ffr.Flame(ex.example_json(name, size=size))
    _, _, cells, cell = fl.layout()
    first, count = sharding.split_chains(chains, world)[rank]
    buf, st, _ = po.oracle_render(fl, count, L, base_seed=4, chain_first=first)
    t = torch.from_numpy(buf.view(np.int64))
    sharding.reduce_buffer(t, cells, cell, dst=0)
    plotted = torch.tensor([st["s_plot"]], dtype=torch.int64)
    dist.reduce(plotted, dst=0)
    if rank == 0:
        q.put((t.numpy().view(np.uint64).copy(), int(plotted.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,size", [("barnsley_fern", [64, 64]), ("tkoz_test3", [64, 36])])
def test_two_rank_reduce_matches_single_process(ffr, po, examples, name, size):
    chains, L, world = 37, 600, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, size, chains, L, q))
             for r in range(world)]
    for p in procs:
        p.start()
    got, plotted = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fl = ffr.Flame(examples.example_json(name, size=size))
    want, st, _ = po.oracle_render(fl, chains, L, base_seed=4)
    _, _, cells, cs = fl.layout()
    gc, gcol = ffr.split_counts_colors(got, cells, cs - 1)
    wc, wcol = ffr.split_counts_colors(want, cells, cs - 1)
    assert np.array_equal(gc, wc)
    assert plotted == st["s_plot"]
    if cs > 1:
        np.testing.assert_allclose(gcol, wcol, rtol=1e-12, atol=1e-12)


def test_chain_ranges_are_disjoint_and_cover(ffr):
    sharding = importlib.import_module("flame-fractal-renderer_b200.sharding")
    for total, world in ((0, 4), (1, 8), (37, 2), (1000, 8), (1001, 8)):
        parts = sharding.split_chains(total, world)
        assert sum(c for _, c in parts) == total
        pos = 0
        for first, count in parts:
            if count:
                assert first == pos
                pos += count
    seen = set()
    for step in range(3):
        for rank in range(4):
            f = sharding.step_chain_range(step, rank, 4, 100)
            r = range(f, f + 100)
            assert not (seen & set(r))
            seen |= set(r)
