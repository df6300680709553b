"""Worker of tests/test_gpu_multi.py::test_torchrun_nccl_reduce_matches_one_gpu (launched with
torch.distributed.run, one process per GPU): every rank renders its contiguous chain range into
torch-owned memory, the buffers are summed into rank 0 with sharding.reduce_buffer over NCCL (one
reduce for counts-only buffers; all-to-all + the library's typed slice sum + sends for buffers with
colour sums), rank 0 compares with the whole job rendered on its own GPU."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ffr = importlib.import_module("flame-fractal-renderer_b200")
    ex = importlib.import_module("flame-fractal-renderer_b200.examples")
    sharding = importlib.import_module("flame-fractal-renderer_b200.sharding")
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ok = True
    for name, size, jit in (("barnsley_fern", [300, 200], ffr.JIT_OFF), ("tkoz_test3", [250, 130], ffr.JIT_ON),
                            ("tkoz_test4", [96, 64], ffr.JIT_OFF)):
        fl = ffr.Flame(ex.example_json(name, size=size))
        _, _, cells, cell = fl.layout()
        chains, L = 4001, 700
        buf = torch.zeros(cells * cell, dtype=torch.int64, device=dev)
        r = ffr.BufferRenderer(fl, devices=[local], external_buffer=buf.data_ptr(),
                               stream=stream.cuda_stream, jit=jit)
        first, count = sharding.split_chains(chains, world)[rank]
        r.render_chains(first, count, L, base_seed=21)
        for _ in range(2):      # twice: the second exchange must not add anything twice
            sharding.reduce_buffer(buf, cells, cell, dst=0, renderer=r)
            if rank != 0:
                buf.zero_()
        torch.cuda.synchronize(dev)
        r.close()
        if rank == 0:
            got = buf.cpu().numpy().view(np.uint64)
            one = ffr.BufferRenderer(fl, devices=[local], jit=jit)
            one.render_chains(0, chains, L, base_seed=21)
            want = one.read_buffer()
            one.close()
            gc, gcol = ffr.split_counts_colors(got, cells, cell - 1)
            wc, wcol = ffr.split_counts_colors(want, cells, cell - 1)
            same = np.array_equal(gc, wc) and (cell == 1 or np.allclose(gcol, wcol, rtol=1e-12, atol=1e-12))
            print("%s: %s (%d plotted)" % (name, "OK" if same else "MISMATCH", int(gc.sum())), flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
