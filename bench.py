#!/usr/bin/env python
"""bench.py -- chaos-game samples/sec into the buffer (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of chains: `chains_per_step` independent
chains of `chain_len` samples per GPU (a whole number of "waves" of resident chain groups),
rendered into the GPU's private buffer. The headline workload (default) is the north-star target
config: csci6360_project at 4096x4096, double/u64, counts only. Weak scaling: every rank
renders its own disjoint chain range (no data-path collective); for N > 1 the timed region ends
with the one exchange step the path has, the sum-reduce of the private buffers to rank 0 over
NCCL/NVLink. Prints ONE JSON line on rank 0.

value     device-timed (CUDA events on the launching stream) throughput of K steps with
          everything resident in HBM.
e2e       the same metric through the reference-facing C ABI with HOST buffers: every step
          uploads a pinned host buffer (the -i resume path), renders and reads the whole buffer
          back to pinned host memory -- H2D and D2H inside the timed region. At N = 1 the steps go
          through the library's streaming interface (ffr_cuda_*_async) on three contexts used in
          turn, so a step's upload, render and read-back overlap its neighbours' other legs; at
          N > 1 through the blocking calls around the NCCL reduce.
roofline  HBM: algorithmic bytes (one RMW of one cell per PLOTTED sample = 2*(1+r)*8 B) per
          launch / average launch duration, against MEASURED_PEAKS.json hbm_gbs.
atomic_roofline  the north star's denominator: bare REDs at the addresses this render scatters
          to (attractor replay, a saturating microbenchmark: the render's scatter ceiling), and
          at uniformly random cells of the same buffer (a reference point, not a ceiling).
cpu_baseline  the unmodified reference (oracle/_ref, BufferRenderer::render) on all host
          cores for a bounded sample of the same workload.
configs   (N = 1, default run) the five BASELINE.json configurations measured the same way, each
          timed >= ~1.2 s with its own clock record; cfg5 (8 GPUs) appears as its per-GPU shard.
strong_scaling  BASELINE config 5 as stated: csci6360_project at 8192^2, a FIXED 1e11 samples
          split over the N ranks, with the buffer reduce and the read-back to the host inside the
          timed region (the weak-scaling headline cannot bend; this curve can).
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
    "barnsley_8192": ("barnsley_fern", [8192, 8192], 8192),
}
DEFAULT_WORKLOAD = "csci6360_4096"
# BASELINE.json "configs", in order
BASELINE_CONFIGS = [("cfg1", "sierpinski_1024"), ("cfg2", "barnsley_2048"), ("cfg3", "tkoz_test3_4096"),
                    ("cfg4", "sierpinski3d_512"), ("cfg5 (per-GPU shard of the 8-GPU job)", "csci6360_8192")]
STRONG_WORKLOAD, STRONG_SAMPLES = "csci6360_8192", 100_000_000_000
METRIC = "chaos-game samples/sec into buffer"
UNIT = "samples/s"

# What bounds each workload's render kernel (with the ncu evidence)
BOUND_NOTE = {
    "csci6360_4096": "K1d is bound by its instruction stream (13 transcendental variations over 5 xforms: "
                     "~1000 mostly dependent instructions per warp-iteration at 5 warps per scheduler, hot "
                     "code as large as the instruction cache; fp64 pipe 37 % busy); the scatter is hidden "
                     "behind the arithmetic and HBM is < 5 % busy",
    "csci6360_8192": "as csci6360_4096; the 512 MiB buffer no longer fits L2 but the scatter stays hidden",
    "tkoz_test3_4096": "K1d, bound by its instruction stream like csci6360; 4 REDs per plotted sample "
                       "(count + 3 colour sums, one 32-byte sector)",
    "sierpinski_1024": "K1e, bound by the L2 atomic units (one RED sector per sample into the "
                       "cell-scrambled L2-resident tile); DRAM idle",
    "barnsley_2048": "K1e, bound by the L2 atomic units (cell-scrambled L2-resident tile); DRAM idle",
    "sierpinski3d_512": "K1e + compact tile of 8x8x8-cell block rows behind a row directory with a "
                        "shared-memory cache: L2 atomic units (the hot rows are L2-resident inside the "
                        "1 GiB buffer) plus the per-sample directory lookup",
    "barnsley_8192": "K1e + compact tile, dense attractor: part of the rows overflow the tile and "
                     "scatter into the 512 MiB buffer directly",
}

# dram__bytes_read.sum + dram__bytes_write.sum per PLOTTED sample of the render kernel, from the
# committed `ncu --set full` captures (profiles/README.md); traffic per launch = this x plotted.
# A workload without a capture reports null.
NCU_DRAM_BYTES_PER_PLOTTED = {
    # (dram read + dram write) / plotted samples of the captured launch (plotted = RED sectors,
    # / 4 for the r = 3 flame whose cell is one sector hit by four REDs)
    "csci6360_4096": ((0.4416e9 + 3.9159e9) / 500.2e6, "profiles/r2_k1d_csci4096.summary.txt"),
    "tkoz_test3_4096": ((3.3952e9 + 12.2317e9) / (2373.8e6 / 4), "profiles/r2_k1d_tkoz3_4096.summary.txt"),
    "sierpinski3d_512": ((7.99e6 + 0.0) / 465.7e6, "profiles/r2_k1e_sierp3d_512_block_rows_dircache.summary.txt"),
    "barnsley_2048": ((25.67e6 + 0.005e6) / 466.0e6, "profiles/r2_k1e_barnsley2048.summary.txt"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed regions (recipe's clocks line).
    One nvidia-smi process for the whole run; mark()/window() cut out a region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.t0 = 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def mark(self):
        """Start of a timed region: only samples taken from here to window() are reported."""
        self.t0 = time.monotonic()

    def window(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.monotonic()
        sm, smax, reasons, power = [], [], set(), []
        window = [(ts, ln) for ts, ln in list(self.lines) if self.t0 <= ts <= t1]
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

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()


def cpu_reference(wl_name, sample, cores=None):
    """The reference's own CPU implementation of the path (BufferRenderer::render through
    oracle/_ref, unmodified sources; the oracle port if _ref is absent) on `cores` host threads."""
    import pyoracle as po
    ex = importlib.import_module("flame-fractal-renderer_b200.examples")
    ename, size, _ = WORKLOADS[wl_name]
    text = ex.example_json(ename, size=size)
    cores = cores or os.cpu_count() or 1
    batch = max(4096, min(1 << 20, (sample + 255) >> 8))  # ffr_buf.cpp:94-101
    if po.have_ref():
        secs, _, _ = po.ref_render_mt(text, sample, cores, batch)
        kind = "reference"
    else:
        ffr = importlib.import_module("flame-fractal-renderer_b200")
        fl = ffr.Flame(text)
        t0 = time.perf_counter()
        po.oracle_render_samples(fl, sample, batch, nthreads=cores)
        secs = time.perf_counter() - t0
        kind = "port"
    return {"value": sample / secs, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d samples of %s, BufferRenderer::render with %d threads, batch %d, %.1f s"
                      % (sample, wl_name, cores, batch, secs)}, secs, batch


def run_reference(args, wl_name):
    """--impl reference: times the reference's CPU implementation on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ename, size, _ = WORKLOADS[wl_name]
    cores = os.cpu_count() or 1
    sample = args.ref_samples
    for _ in range(args.warmup):
        cpu_reference(wl_name, sample, cores)
    total, kind, batch = 0.0, "reference", 0
    for _ in range(args.steps):
        c, secs, batch = cpu_reference(wl_name, sample, cores)
        kind = c["kind"]
        total += secs
    value = sample * args.steps / total
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


class Bench:
    """State shared by the measurements of one process (rank)."""

    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist
        self.np, self.torch, self.dist = np, torch, dist
        self.args = args
        self.ffr = importlib.import_module("flame-fractal-renderer_b200")
        self.ex = importlib.import_module("flame-fractal-renderer_b200.examples")
        self.sharding = importlib.import_module("flame-fractal-renderer_b200.sharding")
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; there is no CPU path to measure")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=self.dev)
        # a dedicated (non-default) stream: handle 0 would mean "library-owned stream" to the ABI
        self.stream = torch.cuda.Stream(self.dev)
        torch.cuda.set_stream(self.stream)
        self.sampler = ClockSampler(self.local_rank)
        if self.rank == 0:
            self.sampler.start()
        self.hbm_peak, self.peak_src = measured_peaks()
        self.jit = self.ffr.JIT_ON if args.jit < 0 else args.jit

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------------------------------
    def measure(self, wl_name, steps, warmup, min_seconds=0.0, ref_samples=0):
        """One workload: device-timed value, roofline, atomic rooflines, e2e, clocks."""
        torch, np, ffr, sharding = self.torch, self.np, self.ffr, self.sharding
        rank, world, dev, stream = self.rank, self.world, self.dev, self.stream
        ename, size, L = WORKLOADS[wl_name]
        flame = ffr.Flame(self.ex.example_json(ename, size=size))
        _, _, cells, cell = flame.layout()
        r_dims = flame.color_dims
        n_elems = cells * cell

        # torch owns the device memory and the stream; the library renders into it
        buf = torch.zeros(n_elems, dtype=torch.int64, device=dev)
        # the flame-specialised kernel is compiled (NVRTC, cached) when the context is created,
        # i.e. outside every timed region, like the ahead-of-time build of the interpreter kernels
        t_create = time.perf_counter()
        rend = ffr.BufferRenderer(flame, devices=[self.local_rank], external_buffer=buf.data_ptr(),
                                  stream=stream.cuda_stream, scatter_mode=self.args.scatter, jit=self.jit)
        t_create = time.perf_counter() - t_create
        jit_info = rend.jit_info
        # one wave = every resident block (SMs x blocks/SM) takes one chain group
        chains_per_step = rend.resident_chains * self.args.waves
        samples_per_step = chains_per_step * L

        def launch_step(step_index):
            # disjoint chain ranges: per step and per rank (weak scaling)
            first = sharding.step_chain_range(step_index, rank, world, chains_per_step)
            rend.render_chains_async(first, chains_per_step, L, base_seed=1)

        # ---- warm-up (>= 3 steps) ----
        warmup = max(warmup, 3)
        launch_step(2_000_000)
        self.barrier()
        t_w = time.perf_counter()
        for w in range(warmup):
            launch_step(1_000_000 + w)
        self.barrier()
        step_s = (time.perf_counter() - t_w) / warmup
        # the exchange step once, untimed: NCCL sets up its peer connections on first use (seconds
        # at N = 8), a one-off like the creation of the context
        sharding.reduce_buffer(buf, cells, cell, dst=0, renderer=rend)
        self.barrier()
        if min_seconds > 0:
            steps = max(steps, int(min_seconds / max(step_s, 1e-4)) + 1)
        buf.zero_()
        st0 = rend.fetch_stats()
        launches0 = rend.launches

        # ---- timed region: K steps (+ the final buffer reduce for N > 1) ----
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev_k = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        self.barrier()
        self.sampler.mark()
        ev0.record(stream)
        ev_k[0].record(stream)
        for k in range(steps):
            launch_step(k)
            ev_k[k + 1].record(stream)
        sharding.reduce_buffer(buf, cells, cell, dst=0, renderer=rend)
        ev1.record(stream)
        self.barrier()
        ms_total = self.max_over_ranks(ev0.elapsed_time(ev1))
        kernel_ms = [ev_k[k].elapsed_time(ev_k[k + 1]) for k in range(steps)]
        clocks = self.sampler.window() if rank == 0 else None
        st1 = rend.fetch_stats()
        launches = rend.launches - launches0
        plotted = st1["s_plot"] - st0["s_plot"]
        iterated = st1["s_iter"] - st0["s_iter"]
        assert iterated == samples_per_step * steps, (iterated, samples_per_step * steps)
        total_samples = samples_per_step * steps * world
        value = total_samples / (ms_total * 1e-3)

        # roofline of the dominant kernel, per launch on this rank
        k_ms = sum(kernel_ms) / len(kernel_ms)
        alg_bytes = plotted / steps * 2 * (1 + r_dims) * 8
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        plotted_per_s = plotted / steps / (k_ms * 1e-3)

        # measured atomic-scatter rooflines on the same buffer, three runs each
        atomic = None
        if rank == 0:
            rend.atomic_roofline(1 << 26)
            rend.atomic_roofline(1 << 26, pattern=1)
            rend.atomic_roofline(1 << 26, pattern=2)
            uni, rep, win = [], [], []
            for _ in range(3):
                ms, n_at = rend.atomic_roofline(1 << 30)
                uni.append(n_at / (ms * 1e-3))
                ms, n_rp = rend.atomic_roofline(1 << 28, pattern=1)
                rep.append(n_rp / (ms * 1e-3))
                ms, n_w = rend.atomic_roofline(1 << 29, pattern=2)
                win.append(n_w / (ms * 1e-3))
            uni.sort()
            rep.sort()
            win.sort()
            ceiling = max(rep[1], win[1])
            atomic = {"ceiling": "attractor replay (better of streamed / windowed)",
                      "attractor_replay_cells_per_s": ceiling,
                      "replay_streamed_runs": rep, "replay_windowed_runs": win,
                      "uniform_random_cells_per_s": uni[1], "uniform_random_runs": uni,
                      "plotted_samples_per_s": plotted_per_s,
                      "frac": plotted_per_s / ceiling,
                      "frac_of_uniform_random": plotted_per_s / uni[1],
                      "note": "replay = bare REDs at the addresses this render scatters to (incl. K1e's "
                              "tile scramble / row directory), 2048 threads/SM, nothing to wait for: the "
                              "scatter ceiling of this render. streamed: the trace is read from memory (an "
                              "L2 sector per 4 REDs); windowed: from shared memory, each window replayed "
                              "many times. uniform random cells of the same buffer are a reference point "
                              "only (an attractor can be L2-resident where uniform addresses stream from HBM)"}
        buf.zero_()
        self.barrier()

        # ---- e2e through the C ABI with host buffers ----
        rend.close()
        nbytes = n_elems * 8
        host_in = torch.zeros(n_elems, dtype=torch.int64).pin_memory()
        host_out = torch.zeros(n_elems, dtype=torch.int64).pin_memory()
        in_np = host_in.numpy().view(np.uint64)
        out_np = host_out.numpy().view(np.uint64)
        ebuf = None
        if world > 1:
            # multi-rank e2e: render into torch memory so NCCL can reduce it, host copies on rank 0
            ebuf = torch.zeros(n_elems, dtype=torch.int64, device=dev)
            e2e_rend = ffr.BufferRenderer(flame, devices=[self.local_rank], external_buffer=ebuf.data_ptr(),
                                          stream=stream.cuda_stream, jit=self.jit)

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
            # three contexts on three streams, used in turn: a step's upload (copy engine), render
            # (SMs) and read-back (copy engine) overlap the neighbouring steps' other legs (the
            # library's streaming interface; every step still uploads its input from pinned host
            # memory and reads its whole result back)
            # The pipeline is only as deep as it pays: a launch that runs for a fifth of a second
            # next to 5 ms of copies gains nothing from overlap and loses a little to three
            # persistent kernels sharing the device (measured on the headline: 0.975 of the device
            # rate with one context, 0.945 with three), so such steps go through one context.
            copy_s = 2 * nbytes / 45e9
            depth = 3 if (copy_s >= 0.05 * step_s or step_s < 0.06) else 1
            e2e_rend = ffr.BufferRenderer(flame, devices=[self.local_rank], jit=self.jit)
            extra = [ffr.BufferRenderer(flame, devices=[self.local_rank], jit=self.jit) for _ in range(depth - 1)]
            extra_host = [torch.zeros(n_elems, dtype=torch.int64).pin_memory() for _ in range(depth - 1)]
            pipe = [(e2e_rend, out_np)] + [(r, h.numpy().view(np.uint64)) for r, h in zip(extra, extra_host)]

            def e2e_step(k):
                r, out = pipe[k % depth]
                r.sync()                      # this context's previous step (`depth` steps ago)
                r.clear_async()
                r.add_buffer_async(in_np)
                first = (k + 500_000) * chains_per_step
                r.render_chains_async(first, chains_per_step, L, base_seed=1)
                r.read_buffer_async(out)

        e2e_depth = 1
        e2e_steps = max(6, min(steps, 20))
        for k in (-3, -2, -1):
            e2e_step(k)
        if world == 1:
            for r, _ in pipe:
                r.sync()
        self.barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            e2e_step(k)
        if world == 1:
            for r, _ in pipe:
                r.sync()
        self.barrier()
        e2e_s = self.max_over_ranks(time.perf_counter() - t0)
        e2e_value = samples_per_step * e2e_steps * world / e2e_s
        if world == 1:
            # the result of the last step really arrived: every sample of a step is iterated once
            done = sum(r.fetch_stats()["s_iter"] for r, _ in pipe)
            assert done == samples_per_step * (e2e_steps + 3), (done, samples_per_step, e2e_steps)
            e2e_depth = depth
            if cell == 1:
                # counts only: each host buffer holds exactly the samples its last step plotted
                assert all(0 < int(o.sum()) <= samples_per_step for _, o in pipe)
            for r in extra:
                r.close()
            del extra_host, pipe
        e2e_rend.close()
        del buf, ebuf, host_in, host_out, in_np, out_np
        torch.cuda.empty_cache()

        cpu = None
        if rank == 0 and world == 1 and ref_samples > 0:
            cpu, _, _ = cpu_reference(wl_name, ref_samples)
        if rank != 0:
            return None
        traffic = None
        traffic_src = None
        if wl_name in NCU_DRAM_BYTES_PER_PLOTTED:
            per, traffic_src = NCU_DRAM_BYTES_PER_PLOTTED[wl_name]
            traffic = per * plotted / steps
        kernel = ("flame-specialised, compiled at context creation (NVRTC %.1f s%s, context %.1f s): "
                  "%d threads x %d blocks/SM, %d chain slots/block, %d registers; %s"
                  % (jit_info["compile_seconds"], ", cached" if jit_info["from_cache"] else "", t_create,
                     jit_info["threads_per_block"], jit_info["blocks_per_sm"], jit_info["slots_per_block"],
                     jit_info["registers"], jit_info["message"].strip())
                  if jit_info["active"] else "ahead-of-time interpreter kernel")
        return {
            "value": value, "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps,
            "timed_seconds": ms_total * 1e-3,
            "config": {"workload": wl_name, "flame": ename, "size": size, "color_dims": r_dims,
                       "chain_len": L, "chains_per_step_per_gpu": chains_per_step,
                       "samples_per_step_per_gpu": samples_per_step, "base_seed": 1,
                       "kernel": kernel,
                       "parallelism": "chain-range sharding x%d, private buffers, final NCCL sum-reduce" % world,
                       "l2": "no flush: the only memory operand is the %.0f MiB accumulation buffer (L2 is "
                             "126 MB), which a render keeps resident across steps; samples are generated "
                             "on device" % (n_elems * 8 / 2**20)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes,
                    "d2h_bytes_per_step": nbytes, "steps": e2e_steps, "contexts_in_flight": e2e_depth},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": self.hbm_peak, "unit": "GB/s",
                         "frac": achieved / self.hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": self.peak_src, "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "plotted_fraction": plotted / iterated,
                         "limiter": BOUND_NOTE.get(wl_name)},
            "atomic_roofline": atomic,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }

    # ------------------------------------------------------------------------------------------
    def strong(self):
        """BASELINE config 5 as stated: a FIXED 1e11 samples of csci6360_project at 8192^2 split over
        the ranks (contiguous chain ranges), private buffers, ONE sum-reduce to rank 0 and the
        read-back of the 512 MiB result to the host -- all inside the timed region."""
        torch, ffr, sharding = self.torch, self.ffr, self.sharding
        ename, size, L = WORKLOADS[STRONG_WORKLOAD]
        flame = ffr.Flame(self.ex.example_json(ename, size=size))
        _, _, cells, cell = flame.layout()
        n_elems = cells * cell
        total = self.args.strong_samples
        chains = (total + L - 1) // L
        first, count = sharding.split_chains(chains, self.world)[self.rank]
        buf = torch.zeros(n_elems, dtype=torch.int64, device=self.dev)
        rend = ffr.BufferRenderer(flame, devices=[self.local_rank], external_buffer=buf.data_ptr(),
                                  stream=self.stream.cuda_stream, jit=self.jit)
        host = torch.zeros(n_elems, dtype=torch.int64).pin_memory()
        rend.render_chains_async(10_000_000_000, rend.resident_chains, L, base_seed=1)   # warm
        self.barrier()
        buf.zero_()
        self.barrier()
        self.sampler.mark()
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(self.stream)
        rend.render_chains_async(first, count, L, base_seed=1)
        sharding.reduce_buffer(buf, cells, cell, dst=0, renderer=rend)
        if self.rank == 0:
            host.copy_(buf, non_blocking=True)
        ev1.record(self.stream)
        self.barrier()
        wall = self.max_over_ranks(time.perf_counter() - t0)
        ms = self.max_over_ranks(ev0.elapsed_time(ev1))
        clocks = self.sampler.window() if self.rank == 0 else None
        plotted = int(host.sum().item()) if self.rank == 0 else 0
        rend.close()
        del buf, host
        torch.cuda.empty_cache()
        if self.rank != 0:
            return None
        return {"workload": STRONG_WORKLOAD, "scaling": "strong", "total_samples": chains * L,
                "n_gpus": self.world, "seconds": ms * 1e-3, "wall_seconds": wall,
                "value": chains * L / (ms * 1e-3), "unit": UNIT, "plotted": plotted,
                "includes": "render of this rank's chain range + buffer sum-reduce to rank 0 + D2H of "
                            "the %.0f MiB result" % (n_elems * 8 / 2**20),
                "clocks": clocks}


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
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the per-config measurements of the five BASELINE configs (N = 1)")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-size cfg5 run")
    ap.add_argument("--strong-samples", type=int, default=STRONG_SAMPLES)
    ap.add_argument("--scatter", type=int, default=0)
    ap.add_argument("--jit", type=int, default=-1,
                    help="run-time compiled flame-specialised kernel: -1 = on for flames with "
                         "non-linear variations, 1 = off (interpreter kernels), 2 = on")
    args = ap.parse_args()
    wl_name = args.workload
    if args.impl == "reference":
        run_reference(args, wl_name)
        return

    b = Bench(args)
    head = b.measure(wl_name, args.steps, args.warmup,
                     ref_samples=0 if (args.no_cpu_baseline or b.world > 1) else args.ref_samples)
    configs = None
    if b.world == 1 and not args.no_configs and wl_name == DEFAULT_WORKLOAD:
        configs = []
        for label, name in BASELINE_CONFIGS:
            m = b.measure(name, 3, 3, min_seconds=1.2,
                          ref_samples=0 if args.no_cpu_baseline else 20_000_000)
            m["baseline_config"] = label
            m["unit"] = UNIT
            configs.append(m)
    strong = None
    if not args.no_strong and wl_name == DEFAULT_WORKLOAD:
        strong = b.strong()
    b.sampler.stop()
    if b.rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": b.world,
            "steps": head["steps"], "warmup": head["warmup"], "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": head["config"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"], "atomic_roofline": head["atomic_roofline"],
            "cpu_baseline": head["cpu_baseline"], "clocks": head["clocks"],
        }
        if configs is not None:
            line["configs"] = configs
        if strong is not None:
            line["strong_scaling"] = strong
        print(json.dumps(line), flush=True)
    if b.world > 1:
        b.dist.destroy_process_group()


if __name__ == "__main__":
    main()
