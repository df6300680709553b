"""Multi-GPU paths on hardware (skipped with fewer than 2 devices): the single-process
multi-device context (chain-range sharding + peer-memory reduce inside libffr_cuda, the
ffr-buf.out --gpus path) must give the same buffer as one device, bit for bit for counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev(ffr):
    return ffr.lib().ffr_cuda_device_count()


@pytest.mark.parametrize("name,size", [("barnsley_fern", [256, 256]), ("tkoz_test3", [160, 90])])
def test_two_devices_equal_one(ffr, po, examples, name, size):
    if _ndev(ffr) < 2:
        pytest.skip("needs 2 GPUs")
    fl = ffr.Flame(examples.example_json(name, size=size))
    _, _, cells, cs = fl.layout()
    r1 = ffr.BufferRenderer(fl, devices=[0])
    assert r1.render(3_000_000, 1000, base_seed=8)
    b1 = r1.read_buffer()
    s1 = r1.stats
    r1.close()
    r2 = ffr.BufferRenderer(fl, devices=[0, 1])
    assert r2.render(3_000_000, 1000, base_seed=8)
    b2 = r2.read_buffer()
    s2 = r2.stats
    # a second read must not add the peer buffer twice
    b2b = r2.read_buffer()
    r2.close()
    assert np.array_equal(b2, b2b)
    c1, col1 = ffr.split_counts_colors(b1, cells, cs - 1)
    c2, col2 = ffr.split_counts_colors(b2, cells, cs - 1)
    if name == "barnsley_fern":
        assert np.array_equal(c1, c2)
        for k in ("s_iter", "s_plot", "xf_dist", "pt_min", "pt_max"):
            assert s1[k] == s2[k]
        want, _, _ = po.oracle_render_samples(fl, 3_000_000, 1000, base_seed=8, nthreads=8)
        assert np.array_equal(b2, want)
    else:
        # same chains, same device code on both GPUs: counts identical, colours to rounding
        assert np.array_equal(c1, c2)
        np.testing.assert_allclose(col1, col2, rtol=1e-12, atol=1e-9)


def test_two_devices_jit_kernel(ffr, examples):
    """The run-time compiled kernel on a two-device context: one module per device, chains handed
    out per device by the kernel's own counter, peer-memory reduce: counts equal one device."""
    if _ndev(ffr) < 2:
        pytest.skip("needs 2 GPUs")
    fl = ffr.Flame(examples.example_json("csci6360_project", size=[160, 90]))
    _, _, cells, cs = fl.layout()
    bufs = []
    for devs, jit in (([0], ffr.JIT_OFF), ([0, 1], ffr.JIT_ON)):
        r = ffr.BufferRenderer(fl, devices=devs, jit=jit)
        assert r.render(3_000_000, 1000, base_seed=8)
        assert bool(r.jit_info["active"]) == (jit == ffr.JIT_ON)
        bufs.append((r.read_buffer(), r.stats))
        r.close()
    assert np.array_equal(bufs[0][0], bufs[1][0])
    for k in ("s_iter", "s_plot", "xf_dist", "pt_min", "pt_max"):
        assert bufs[0][1][k] == bufs[1][1][k]


def test_add_buffer_resume(ffr, po, examples):
    """-i semantics (buffer_renderer.hpp:375-452): counts add as integers, colours as floats."""
    fl = ffr.Flame(examples.example_json("tkoz_test3", size=[96, 54]))
    _, _, cells, cs = fl.layout()
    r = ffr.BufferRenderer(fl)
    r.render_chains(0, 300, 512, base_seed=3)
    first = r.read_buffer().copy()
    r.clear()
    r.add_buffer(first)
    r.add_buffer(first)
    twice = r.read_buffer()
    r.close()
    c1, col1 = ffr.split_counts_colors(first, cells, cs - 1)
    c2, col2 = ffr.split_counts_colors(twice, cells, cs - 1)
    assert np.array_equal(c2, 2 * c1)
    np.testing.assert_array_equal(col2, col1 + col1)
    assert r.bytes == first.nbytes


def test_segmented_render_equals_single_launch(ffr, examples):
    """ffr_cuda_render with a progress callback cuts a large render into several launches (each
    at least four waves of the resident chains); which launch runs a chain changes nothing:
    counts and statistics equal those of the single launch without a callback."""
    fl = ffr.Flame(examples.example_json("sierpinski_triangle", size=[128, 128]))
    r = ffr.BufferRenderer(fl, jit=ffr.JIT_OFF)
    L = 256
    samples = (r.resident_chains * 9 + 7) * L + 31      # > 2 segments, ragged last chain
    calls = []
    assert r.render(samples, L, base_seed=3, progress=lambda d, t: calls.append((d, t)))
    seg_buf, seg_st = r.read_buffer(), r.stats
    r.close()
    assert len(calls) >= 2 and calls[-1][0] == calls[-1][1] == (samples + L - 1) // L
    assert all(a[0] < b[0] for a, b in zip(calls, calls[1:]))
    r = ffr.BufferRenderer(fl, jit=ffr.JIT_OFF)
    assert r.render(samples, L, base_seed=3)
    one_buf, one_st = r.read_buffer(), r.stats
    r.close()
    assert np.array_equal(seg_buf, one_buf)
    for k in ("s_iter", "s_plot", "xf_dist", "pt_min", "pt_max", "n_bad"):
        assert seg_st[k] == one_st[k], k
    assert seg_st["s_iter"] == samples


def test_histogram_sum_max_and_progress(ffr, examples):
    fl = ffr.Flame(examples.example_json("sierpinski_triangle", size=[128, 128]))
    r = ffr.BufferRenderer(fl)
    calls = []
    assert r.render(2_000_000, 512, base_seed=1, progress=lambda d, t: calls.append((d, t)))
    buf = r.read_buffer()
    s, m = r.histogram_sum_max()
    assert s == int(buf.sum()) == 2_000_000 == r.stats["s_plot"]
    assert m == int(buf.max())
    assert calls and calls[-1][0] == calls[-1][1] == (2_000_000 + 511) // 512
    # render() argument checks of the reference (buffer_renderer.hpp:279-285)
    with pytest.raises(ffr.FfrError, match="batch size too small"):
        r.render(1000, 255)
    assert r.render(0, 4096)
    r.close()


def test_ffr_buf_cli_roundtrip(ffr, po, examples, tmp_path):
    """ffr-buf.out: same flags and buffer file format; -i adds the previous output."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(ffr.LIB_PATH), "ffr-buf.out")
    flame = tmp_path / "fern.json"
    flame.write_text(examples.example_json("barnsley_fern", size=[200, 100]))
    out1 = tmp_path / "a.buf"
    out2 = tmp_path / "b.buf"
    p = subprocess.run([exe, "-f", str(flame), "-o", str(out1), "-s", "1500000", "-b", "1000",
                        "--seed", "5", "-t", "3"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "samples plotted: 1500000" in p.stderr and "samples/sec" in p.stderr
    a = np.fromfile(out1, dtype=np.uint64)
    fl = ffr.Flame(flame.read_text())
    want, _, _ = po.oracle_render_samples(fl, 1_500_000, 1000, base_seed=5, nthreads=8)
    assert np.array_equal(a, want)
    p = subprocess.run([exe, "-f", str(flame), "-i", str(out1), "-i", str(out1), "-o", str(out2),
                        "-s", "0"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "(not rendering)" in p.stderr
    assert np.array_equal(np.fromfile(out2, dtype=np.uint64), 2 * a)


def test_ffr_buf_cli_jit_flag(ffr, examples, tmp_path):
    """--jit / --no-jit give byte-identical buffer files (same chains, same arithmetic)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(ffr.LIB_PATH), "ffr-buf.out")
    flame = tmp_path / "csci.json"
    flame.write_text(examples.example_json("csci6360_project", size=[192, 108]))
    outs = []
    for flag in ("--jit", "--no-jit"):
        out = tmp_path / (flag.strip("-") + ".buf")
        p = subprocess.run([exe, "-f", str(flame), "-o", str(out), "-s", "2000000", "-b", "1000",
                            "--seed", "9", flag], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        outs.append(np.fromfile(out, dtype=np.uint64))
    assert outs[0].sum() > 0 and np.array_equal(outs[0], outs[1])


def test_atomic_replay_with_jit_kernel(ffr, examples):
    """The attractor replay records its trace through whichever render kernel is active."""
    fl = ffr.Flame(examples.example_json("csci6360_project", size=[256, 256]))
    r = ffr.BufferRenderer(fl, jit=ffr.JIT_ON)
    r.render_chains(0, 512, 256, base_seed=1)
    before = r.fetch_stats()
    ms1, n1 = r.atomic_roofline(1 << 22, pattern=1)
    assert ms1 > 0 and 0 < n1 <= (1 << 22) + r.resident_chains
    assert r.fetch_stats() == before
    r.close()


def test_atomic_roofline_patterns(ffr, examples):
    """The scatter microbenchmarks: uniform cells and attractor replay both run,
    report how many cells they hit, and leave the render statistics untouched."""
    fl = ffr.Flame(examples.example_json("barnsley_fern", size=[256, 256]))
    r = ffr.BufferRenderer(fl)
    r.render_chains(0, 512, 256, base_seed=1)
    before = r.fetch_stats()
    ms0, n0 = r.atomic_roofline(1 << 22, pattern=0)
    ms1, n1 = r.atomic_roofline(1 << 22, pattern=1)
    ms2, n2 = r.atomic_roofline(1 << 24, pattern=2)
    assert ms0 > 0 and ms1 > 0 and ms2 > 0
    assert n0 >= (1 << 22) * 0.5
    # the fern plots every sample, so the streamed replay hits one cell per recorded sample:
    # resident_chains chains x ceil(n / chains) samples
    chains = r.resident_chains & ~1
    assert n1 == chains * -(-(1 << 22) // chains)
    # the windowed replay issues windows x repetitions REDs, about what was asked for
    assert (1 << 23) <= n2 <= (1 << 24)
    assert r.fetch_stats() == before
    r.close()


def test_streaming_interface_equals_blocking_calls(ffr, po, examples):
    """ffr_cuda_{clear,add_buffer,read_buffer}_async + render_chains_async on page-locked
    buffers give the buffer the blocking calls give; pageable host memory is refused; the
    blocking add gives the same result from pinned and from pageable host memory."""
    import torch
    fl = ffr.Flame(examples.example_json("tkoz_test3", size=[320, 200]))   # mixed u64 / f64 cells
    _, _, cells, cs = fl.layout()
    prev, _, _ = po.oracle_render(fl, 64, 500, base_seed=8, nthreads=4)
    pinned_in = torch.from_numpy(prev.view(np.int64)).pin_memory()
    pinned_out = torch.zeros(cells * cs, dtype=torch.int64).pin_memory()
    in_np = pinned_in.numpy().view(np.uint64)
    out_np = pinned_out.numpy().view(np.uint64)
    r = ffr.BufferRenderer(fl)
    for _ in range(2):          # twice: the second pass starts from a cleared buffer again
        r.clear_async()
        r.add_buffer_async(in_np)
        r.render_chains_async(0, 300, 700, base_seed=5)
        r.read_buffer_async(out_np)
        r.sync()
    got = out_np.copy()
    with pytest.raises(ffr.FfrError, match="page-locked"):
        r.add_buffer_async(prev)
    with pytest.raises(ffr.FfrError, match="page-locked"):
        r.read_buffer_async(np.zeros(cells * cs, dtype=np.uint64))
    r.close()
    want = []
    for src in (in_np, prev):    # pinned and pageable host memory
        b = ffr.BufferRenderer(fl)
        b.add_buffer(src)
        assert b.render_chains(0, 300, 700, base_seed=5)
        want.append(b.read_buffer())
        b.close()
    gc, gcol = ffr.split_counts_colors(got, cells, cs - 1)
    for w in want:
        wc, wcol = ffr.split_counts_colors(w, cells, cs - 1)
        assert np.array_equal(gc, wc)
        np.testing.assert_allclose(gcol, wcol, rtol=1e-12, atol=1e-12)
    # and the -i semantics hold: counts = previous counts + this render's
    pc, _ = ffr.split_counts_colors(prev, cells, cs - 1)
    assert int(gc.sum()) == int(pc.sum()) + 0 + int((gc - pc).sum()) and (gc >= pc).all()


def _run_cli(ffr, args, env=None):
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(ffr.LIB_PATH), "ffr-buf.out")
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe] + args, capture_output=True, text=True, env=e, timeout=600)


@pytest.mark.parametrize("name,size,flag", [("barnsley_fern", [256, 160], "--no-jit"),
                                            ("tkoz_test3", [192, 108], "--jit"),
                                            ("csci6360_project", [192, 108], "--jit")])
def test_cli_one_process_per_gpu_equals_one_gpu(ffr, examples, tmp_path, name, size, flag):
    """ffr-buf.out --gpus 2, both forms: one process with a two-device context (default), and
    FFR_MULTI_PROCESS=1: two processes (fork before any CUDA call), the worker's buffer added to
    the collector's over CUDA IPC peer memory. Counts byte-identical to the 1-GPU file, the same
    stderr report (s_iter, s_plot, xform selection)."""
    if ffr.lib().ffr_cuda_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    flame = tmp_path / "f.json"
    flame.write_text(examples.example_json(name, size=size))
    fl = ffr.Flame(flame.read_text())
    _, _, cells, cs = fl.layout()
    common = ["-f", str(flame), "-s", "3000000", "-b", "1000", "--seed", "11", flag]
    outs, reports = [], []
    for tag, extra, env in (("one", [], None), ("two", ["--gpus", "2"], None),
                            ("two_mp", ["--gpus", "2"], {"FFR_MULTI_PROCESS": "1"})):
        out = tmp_path / (tag + ".buf")
        p = _run_cli(ffr, common + ["-o", str(out)] + extra, env)
        assert p.returncode == 0, p.stderr
        outs.append(np.fromfile(out, dtype=np.uint64))
        reports.append([ln for ln in p.stderr.split("\n")
                        if ln.startswith(("samples iterated", "samples plotted", "xform selection",
                                          "extreme coordinates"))])
    c0, col0 = ffr.split_counts_colors(outs[0], cells, cs - 1)
    for o in outs[1:]:
        c, col = ffr.split_counts_colors(o, cells, cs - 1)
        assert np.array_equal(c, c0)
        if cs > 1:
            np.testing.assert_allclose(col, col0, rtol=1e-12, atol=1e-12)
    assert reports[0] == reports[1] == reports[2] and len(reports[0]) == 4


def test_cli_worker_failure_is_reported(ffr, examples, tmp_path):
    """More processes than devices: the worker without a device reports through its pipe, the
    collector prints the error and exits non-zero (no hang, no partial output file)."""
    n = ffr.lib().ffr_cuda_device_count()
    flame = tmp_path / "f.json"
    flame.write_text(examples.example_json("barnsley_fern", size=[64, 64]))
    out = tmp_path / "x.buf"
    p = _run_cli(ffr, ["-f", str(flame), "-o", str(out), "-s", "1000000", "-b", "1000", "--seed", "1",
                       "--gpus", str(n + 1)], {"FFR_MULTI_PROCESS": "1"})
    assert p.returncode == 1
    assert "ERROR: worker %d" % n in p.stderr and "device index out of range" in p.stderr
    assert not out.exists()
    # the default form (one process) reports the same condition itself
    p = _run_cli(ffr, ["-f", str(flame), "-o", str(out), "-s", "1000000", "-b", "1000", "--seed", "1",
                       "--gpus", str(n + 1)])
    assert p.returncode == 1 and "device index out of range" in p.stderr and not out.exists()


def test_torchrun_nccl_reduce_matches_one_gpu(ffr):
    """One process per GPU over NCCL (the bench's multi-GPU host): tests/nccl_reduce_worker.py on
    two ranks; counts bit-identical to one GPU, colour sums to rounding, for a counts-only buffer
    (one NCCL reduce) and for buffers with colour sums (all-to-all + K2d typed slice sum + sends)."""
    import os
    import subprocess
    import sys
    if ffr.lib().ffr_cuda_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577",
                        os.path.join(here, "nccl_reduce_worker.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert p.stdout.count(": OK") == 3, p.stdout
