#!/usr/bin/env python
"""bench.py -- chaos-game samples/sec into the buffer (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of chains: `chains_per_step` independent
chains of `chain_len` samples per GPU (a whole number of "waves" of resident chain groups),
rendered into the GPU's private buffer. The workload (default) is the north-star target config:
csci6360_project at 4096x4096, double/u64, counts only. Weak scaling: every rank renders its
own disjoint chain range (no data-path collective); for N > 1 the timed region ends with the
one exchange step the path has, the sum-reduce of the private buffers to rank 0 over
NCCL/NVLink. Prints ONE JSON line on rank 0.

value     device-timed (CUDA events on the launching stream) throughput of K steps with
          everything resident in HBM.
e2e       the same metric through the reference-facing C ABI with HOST buffers: every step
          uploads the previous host buffer (the -i resume path, ffr_cuda_add_buffer), renders
          (blocking ffr_cuda_render_chains incl. statistics read-back) and reads the whole
          buffer back (ffr_cuda_read_buffer) -- H2D and D2H inside the timed region.
roofline  HBM: algorithmic bytes (one RMW of one cell per PLOTTED sample = 2*(1+r)*8 B) per
          launch / average launch duration, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline  the unmodified reference (oracle/_ref, BufferRenderer::render) on all host
          cores for a bounded sample of the same workload.
"""

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # name: (example flame, size, chain_len)
    "csci6360_4096": ("csci6360_project", [4096, 4096], 8192),
    "csci6360_8192": ("csci6360_project", [8192, 8192], 8192),
    "sierpinski_1024": ("sierpinski_triangle", [1024, 1024], 8192),
    "barnsley_2048": ("barnsley_fern", [2048, 2048], 8192),
    "tkoz_test3_4096": ("tkoz_test3", [4096, 4096], 8192),
    "sierpinski3d_512": ("sierpinski_triangle_3d", [512, 512, 512], 8192),
}
DEFAULT_WORKLOAD = "csci6360_4096"
METRIC = "chaos-game samples/sec into buffer"
UNIT = "samples/s"


# dram__bytes_read.sum + dram__bytes_write.sum per PLOTTED sample of the render kernel, from the
# committed `ncu --set full` captures (profiles/README.md); traffic per launch = this x plotted
NCU_DRAM_BYTES_PER_PLOTTED = {
    "csci6360_4096": (0.4210e9 + 3.8049e9) / 500.0e6,   # profiles/r1_k1d_csci4096_sincos (500.0e6 plotted = RED sectors)
    "tkoz_test3_4096": (1.8337e9 + 6.6085e9) / 327.7e6,  # profiles/r1_k1d_tkoz3_4096
    "sierpinski3d_512": (8.99e6 + 0.006e6) / 1862.3e6,  # profiles/r1_k1e_sierp3d_512_compact_tile
    "barnsley_2048": (25.83e6 + 0.008e6) / 1862.3e6,    # profiles/r1_k1e_barnsley2048
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def mark(self):
        """Start of the timed region: only samples taken from here to stop() are reported
        (the sampler itself is started before the warm-up so that nvidia-smi's start-up time
        does not eat the region)."""
        self.t0 = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.monotonic()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        t0 = getattr(self, "t0", 0.0)
        window = [(ts, ln) for ts, ln in list(self.lines) if t0 <= ts <= t1]
        if not window:      # region shorter than one sampling period: the nearest samples
            window = list(self.lines)[-2:]
        for ts, ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, wl_name):
    """--impl reference: the reference's own CPU implementation of the path
    (BufferRenderer::render through oracle/_ref, unmodified sources) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pyoracle as po
    ex = importlib.import_module("flame-fractal-renderer_b200.examples")
    ename, size, _ = WORKLOADS[wl_name]
    text = ex.example_json(ename, size=size)
    cores = os.cpu_count() or 1
    have_ref = po.have_ref()
    sample = args.ref_samples
    batch = max(4096, min(1 << 20, (sample + 255) >> 8))  # ffr_buf.cpp:94-101

    def one_step():
        if have_ref:
            secs, st, _ = po.ref_render_mt(text, sample, cores, batch)
            return secs
        ffr = importlib.import_module("flame-fractal-renderer_b200")
        fl = ffr.Flame(text)
        t0 = time.perf_counter()
        po.oracle_render_samples(fl, sample, batch, nthreads=cores)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    t = [one_step() for _ in range(args.steps)]
    total = sum(t)
    value = sample * args.steps / total
    kind = "reference" if have_ref else "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, "flame": ename, "size": size,
                   "samples_per_step": sample, "batch_size": batch, "threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d samples per step of %s via BufferRenderer::render, %d threads"
                                   % (sample, wl_name, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--waves", type=int, default=4,
                    help="chain groups per resident block and step")
    ap.add_argument("--ref-samples", type=int, default=100_000_000,
                    help="bounded CPU sample per step for the reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scatter", type=int, default=0)
    ap.add_argument("--jit", type=int, default=-1,
                    help="run-time compiled flame-specialised kernel: -1 = on for flames with "
                         "non-linear variations, 1 = off (interpreter kernels), 2 = on")
    args = ap.parse_args()
    wl_name = args.workload
    if args.impl == "reference":
        run_reference(args, wl_name)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    ffr = importlib.import_module("flame-fractal-renderer_b200")
    ex = importlib.import_module("flame-fractal-renderer_b200.examples")
    sharding = importlib.import_module("flame-fractal-renderer_b200.sharding")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; there is no CPU path to measure")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    ename, size, L = WORKLOADS[wl_name]
    flame = ffr.Flame(ex.example_json(ename, size=size))
    _, _, cells, cell = flame.layout()
    r_dims = flame.color_dims
    n_elems = cells * cell

    # torch owns the device memory and the stream; the library renders into it
    buf = torch.zeros(n_elems, dtype=torch.int64, device=dev)
    # a dedicated (non-default) stream: handle 0 would mean "library-owned stream" to the ABI
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    # the flame-specialised kernel is compiled (NVRTC, cached) when the context is created, i.e.
    # outside every timed region, like the ahead-of-time build of the interpreter kernels
    jit = args.jit
    if jit < 0:
        jit = ffr.JIT_ON   # flame-specialised kernels: K1d (variations), K1e (pure-affine flames)
    t_create = time.perf_counter()
    rend = ffr.BufferRenderer(flame, devices=[local_rank], external_buffer=buf.data_ptr(),
                              stream=stream.cuda_stream, scatter_mode=args.scatter, jit=jit)
    t_create = time.perf_counter() - t_create
    jit_info = rend.jit_info
    # one wave = every resident block (SMs x blocks/SM) takes one chain group of 256 chains
    chains_per_step = rend.resident_chains * args.waves
    samples_per_step = chains_per_step * L

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def launch_step(step_index):
        # disjoint chain ranges: per step and per rank (weak scaling)
        first = sharding.step_chain_range(step_index, rank, world, chains_per_step)
        rend.render_chains_async(first, chains_per_step, L, base_seed=1)

    def reduce_to_rank0():
        sharding.reduce_buffer(buf, cells, cell, dst=0, renderer=rend)

    # ---- warm-up (>= 3 steps); the clock sampler starts here, reports the timed region only ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for w in range(max(args.warmup, 0)):
        launch_step(1_000_000 + w)
    barrier()
    buf.zero_()
    st0 = rend.fetch_stats()
    launches0 = rend.launches

    # ---- timed region: K steps (+ the final buffer reduce for N > 1) ----
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev_k = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    sampler.mark()
    ev0.record(stream)
    ev_k[0].record(stream)
    for k in range(args.steps):
        launch_step(k)
        ev_k[k + 1].record(stream)
    reduce_to_rank0()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = [ev_k[k].elapsed_time(ev_k[k + 1]) for k in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    st1 = rend.fetch_stats()
    launches = rend.launches - launches0
    plotted = st1["s_plot"] - st0["s_plot"]
    iterated = st1["s_iter"] - st0["s_iter"]
    assert iterated == samples_per_step * args.steps, (iterated, samples_per_step * args.steps)
    total_samples = samples_per_step * args.steps * world
    value = total_samples / (ms_total * 1e-3)

    # roofline of the dominant (only) kernel, per launch on this rank
    hbm_peak, peak_src = measured_peaks()
    k_ms = sum(kernel_ms) / len(kernel_ms)
    alg_bytes = plotted / args.steps * 2 * (1 + r_dims) * 8
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9

    # measured atomic-scatter rooflines on the same buffer: uniformly random cells,
    # and a replay of this flame's own attractor trace (same warps, same hot-cell collisions)
    rend.atomic_roofline(1 << 28)  # warm
    atomic_ms, n_at = rend.atomic_roofline(1 << 30)
    atomics_per_s = n_at / (atomic_ms * 1e-3)
    replay_ms, n_rp = rend.atomic_roofline(1 << 28, pattern=1)
    replay_per_s = n_rp / (replay_ms * 1e-3)
    buf.zero_()

    # ---- e2e through the C ABI with host buffers ----
    rend.close()
    e2e_rend = ffr.BufferRenderer(flame, devices=[local_rank], jit=jit)
    nbytes = n_elems * 8
    host_in = torch.zeros(n_elems, dtype=torch.int64).pin_memory()
    host_out = torch.zeros(n_elems, dtype=torch.int64).pin_memory()
    in_np = host_in.numpy().view(np.uint64)
    out_np = host_out.numpy().view(np.uint64)

    if world > 1:
        # multi-rank e2e: render into torch memory so NCCL can reduce it, host copies on rank 0
        e2e_rend.close()
        ebuf = torch.zeros(n_elems, dtype=torch.int64, device=dev)
        e2e_rend = ffr.BufferRenderer(flame, devices=[local_rank], external_buffer=ebuf.data_ptr(),
                                      stream=stream.cuda_stream, jit=jit)

        def e2e_step(k):
            ebuf.zero_()
            if rank == 0:
                e2e_rend.add_buffer(in_np)
            first = sharding.step_chain_range(k + 500_000, rank, world, chains_per_step)
            e2e_rend.render_chains(first, chains_per_step, L, base_seed=1)
            sharding.reduce_buffer(ebuf, cells, cell, dst=0, renderer=e2e_rend)
            if rank == 0:
                e2e_rend.read_buffer(out_np)
    else:
        def e2e_step(k):
            e2e_rend.clear()
            e2e_rend.add_buffer(in_np)
            first = (k + 500_000) * chains_per_step
            e2e_rend.render_chains(first, chains_per_step, L, base_seed=1)
            e2e_rend.read_buffer(out_np)

    e2e_step(-1)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_step(k)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_samples / e2e_s
    e2e_rend.close()

    # ---- CPU baseline: the unmodified reference on the host cores (rank 0, N == 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import pyoracle as po
        cores = os.cpu_count() or 1
        text = ex.example_json(ename, size=size)
        sample = args.ref_samples
        batch = max(4096, min(1 << 20, (sample + 255) >> 8))
        if po.have_ref():
            secs, _, _ = po.ref_render_mt(text, sample, cores, batch)
            kind = "reference"
        else:
            t0 = time.perf_counter()
            po.oracle_render_samples(flame, sample, batch, nthreads=cores)
            secs = time.perf_counter() - t0
            kind = "port"
        cpu = {"value": sample / secs, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d samples of %s, BufferRenderer::render with %d threads, batch %d, %.1f s"
                         % (sample, wl_name, cores, batch, secs)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "flame": ename, "size": size, "color_dims": r_dims,
                       "chain_len": L, "chains_per_step_per_gpu": chains_per_step,
                       "samples_per_step_per_gpu": samples_per_step, "base_seed": 1,
                       "kernel": ("flame-specialised, compiled at context creation (NVRTC %.1f s%s, "
                                  "context %.1f s): %d threads x %d blocks/SM, %d chain slots/block, "
                                  "%d registers; %s" % (jit_info["compile_seconds"],
                                                    ", cached" if jit_info["from_cache"] else "",
                                                    t_create, jit_info["threads_per_block"],
                                                    jit_info["blocks_per_sm"], jit_info["slots_per_block"],
                                                    jit_info["registers"], jit_info["message"].strip())
                                  if jit_info["active"] else "ahead-of-time interpreter kernel"),
                       "parallelism": "chain-range sharding x%d, private buffers, "
                                      "final NCCL sum-reduce" % world,
                       "l2": "no flush: the only memory operand is the %.0f MiB accumulation "
                             "buffer (L2 is 126 MB), which a render keeps resident across steps; "
                             "samples are generated on device" % (n_elems * 8 / 2**20)},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak,
                         "traffic": (NCU_DRAM_BYTES_PER_PLOTTED[wl_name] * plotted / args.steps
                                     if wl_name in NCU_DRAM_BYTES_PER_PLOTTED else None),
                         "peak_source": peak_src, "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "plotted_fraction": plotted / iterated,
                         "note": "fp64-issue bound on this flame, not HBM bound: see DESIGN.md"},
            "atomic_roofline": {"uniform_random_cells_per_s": atomics_per_s,
                                "attractor_replay_cells_per_s": replay_per_s,
                                "plotted_samples_per_s": plotted / args.steps / (k_ms * 1e-3),
                                "frac": (plotted / args.steps / (k_ms * 1e-3)) / atomics_per_s,
                                "frac_of_replay": (plotted / args.steps / (k_ms * 1e-3)) / replay_per_s},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
