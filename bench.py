#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload pconv|rfft]

Headline workload (BASELINE.json configs[4], "metric" second half): partitioned convolution with 10 s
impulse responses (480,000 taps -> 937 partitions of 512 samples), 1024 channels per GPU, one distinct
IR per channel. One "step" = one 512-sample streaming block for every channel (one fused kernel launch).
  value = real-time channels @ 48 kHz = channels * (512 / 48000 s) / t_step          (whole job, all GPUs)
The other half of the metric (batched 1024 x 65536-point real FFT, GB/s) is measured in the same run and reported
as the "fft" object of the line -- value, roofline (achieved / peak / frac / traffic), end-to-end through the host API
and, at N = 1, its own cpu_baseline; `--workload rfft` makes it the headline instead. "secondary" (N = 1) carries the
batched 1024-point complex FFT, the config-2 round trip, the host-API latencies and the direct convolution of config 4.

Multi-GPU (torchrun, one rank per GPU): channels are sharded, every rank owns all state of its channels,
there is no data-path collective; NCCL carries only the barrier and the max-over-ranks of the step time.
The headline scaling is weak (1024 channels per GPU, 1024 transforms per GPU); the "strong" object is the literal
BASELINE configs[4] split -- 1024 channels and 1024 transforms in total, 1024/N per GPU. "e2e_single_process" is the
same weak workload driven by ONE process through the library's multi-GPU handle (devices=[0..N-1]: one worker thread
and stream per device, include/b200fft.h), measured by rank 0 while the other ranks wait.

`--impl reference` times the reference's own implementation on the host CPU cores: the unmodified
reference sources compiled against oracle/minicl when oracle/_ref/libclfft_ref.so is present
(kind "reference"), else the C restatement (kind "port"); one reference object per channel, one channel
per thread. That leg and the `cpu_baseline` object are the only places the oracle is executed here.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000.0
PTS = 512
CVS = 480000  # 10 s at 48 kHz -> 937 partitions (truncating)
CHANNELS_PER_GPU = 1024
RFFT_SIZE = 65536
RFFT_BATCH = 1024
RFFT_KERNEL = "fft_sm_kernel<false,kSmRealFwd> (one SM per transform, one HBM pass, split fused)"
RFFT_KERNEL_SMALL = "large_cols_kernel<7,8> + large_rows_kernel<7,8,REAL> (four-step pair, < 96 transforms per call)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def ncu_traffic(key):
    """per-launch DRAM bytes of the dominant kernel from the committed ncu capture, or None"""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(key)
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def __enter__(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # a CPU-only rendezvous for the phase in which rank 0 alone drives every GPU through the library's multi-GPU
        # handle: an NCCL barrier would park a spinning kernel on the other ranks' GPUs for its whole duration
        global _CPU_GROUP
        _CPU_GROUP = dist.new_group(backend="gloo")
    return rank, world, local


_CPU_GROUP = None


def cpu_barrier():
    import torch.distributed as dist

    if _CPU_GROUP is not None:
        dist.barrier(group=_CPU_GROUP)


def barrier():
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()


def event_time_ms(fn, steps, warmup):
    """W untimed calls, then exactly K calls bracketed by barrier+synchronize, timed with CUDA events on the
    launching (current) stream. Returns total ms of this rank."""
    import torch

    for i in range(warmup):
        fn(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def synth_ir_dev(channels, cvs, seed):
    """decaying-noise IRs (-60 dB over the IR length), L2-normalised, generated on the device"""
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    n = torch.arange(cvs, device="cuda", dtype=torch.float32)
    env = torch.exp(-6.9078 * n / cvs)
    ir = torch.randn(channels, cvs, generator=g, device="cuda") * env
    ir /= ir.norm(dim=1, keepdim=True)
    return ir.contiguous()


def bench_pconv(eng, local, rank, world, steps, warmup, channels=CHANNELS_PER_GPU, cvs=CVS, pts=PTS, e2e=True,
                spot_check=False):
    import numpy as np
    import torch

    from opencl_fft_b200.shard import max_over_ranks

    conv = eng.Clpconv(local, cvs, pts, channels=channels)
    if conv.get_cl_err():
        raise RuntimeError("Clpconv: " + eng.cl_error_string(conv.get_cl_err()) + " " + eng.last_cuda_error())
    ir = synth_ir_dev(channels, cvs, 7000 + rank)
    assert conv.push_ir_dev(ir, cvs) == 0
    torch.cuda.synchronize()
    del ir
    nring = 8
    g = torch.Generator(device="cuda").manual_seed(3000 + rank)
    x = (torch.rand(nring, channels, pts, generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty(channels, pts, device="cuda")

    def step(i):
        rc = conv.convolution_dev(y, x[i % nring])
        assert rc == 0, rc

    with ClockSampler(local) as clk:
        ms = event_time_ms(step, steps, warmup)
    ms_max = max_over_ranks(ms)
    nparts = conv.nparts
    bytes_per_launch = channels * 8 * pts * (2 * nparts + 3)
    res = {
        "ms_per_step": ms_max / steps,
        "ms_per_step_rank": ms / steps,
        "channels_per_gpu": channels,
        "nparts": nparts,
        "bytes_per_launch": bytes_per_launch,
        "clocks": clk.summary(),
        "launches": steps,
    }
    # end to end through the reference-facing host API: pinned host buffers in, pinned host buffers out,
    # H2D + kernel + D2H inside every timed call (the call is synchronous like the reference's)
    if e2e:
        hx = torch.empty(nring, channels, pts).pin_memory()
        hx.copy_(x.cpu())
        hy = torch.empty(channels, pts).pin_memory()
        hx_np, hy_np = hx.numpy(), hy.numpy()
        conv.reset()
        for i in range(warmup):
            assert conv.convolution(hy_np, hx_np[i % nring]) == 0
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            assert conv.convolution(hy_np, hx_np[(warmup + i) % nring]) == 0
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        e2e_ms = max_over_ranks((t1 - t0) * 1e3)
        res["e2e_ms_per_step"] = e2e_ms / steps
        res["h2d_bytes_per_step"] = channels * pts * 4
        res["d2h_bytes_per_step"] = channels * pts * 4
        # the host path and the device path run the same kernel on the same state: same bits
        conv.reset()
        conv2_out = torch.empty(channels, pts, device="cuda")
        assert conv.convolution(hy_np, hx_np[0]) == 0
        conv.reset()
        assert conv.convolution_dev(conv2_out, x[0]) == 0
        torch.cuda.synchronize()
        res["e2e_matches_device_path"] = bool(np.array_equal(conv2_out.cpu().numpy(), hy_np))
        # the reference's API takes any pointer (cl_conv.cpp:399,455): the same call on pageable numpy buffers
        px, py = np.array(hx_np[0]), np.empty((channels, pts), np.float32)
        conv.reset()
        for i in range(2):
            assert conv.convolution(py, px) == 0
        barrier()
        n = max(3, min(steps, 20))
        t0 = time.perf_counter()
        for i in range(n):
            assert conv.convolution(py, px) == 0
        t1 = time.perf_counter()
        barrier()
        res["e2e_pageable_ms_per_step"] = max_over_ranks((t1 - t0) * 1e3) / n
    if spot_check:
        res["parity_spot_check"] = pconv_spot_check(conv, channels, cvs, pts, 7000 + rank)
    conv.close()
    return res


def pconv_spot_check(conv, channels, cvs, pts, seed):
    """Outside every timed region: two blocks of two channels of the benched handle against the CPU oracle
    (relative L2; the parity contract is 1e-5)."""
    import numpy as np
    import torch

    import oracle

    P = oracle.port()
    ir = synth_ir_dev(channels, cvs, seed)
    picks = [0, channels - 1]
    irs = {k: ir[k].cpu().numpy() for k in picks}
    del ir
    conv.reset()
    g = torch.Generator(device="cuda").manual_seed(99)
    x = (torch.rand(2, channels, pts, generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty(2, channels, pts, device="cuda")
    for t in range(2):
        assert conv.convolution_dev(y[t], x[t]) == 0
    torch.cuda.synchronize()
    worst = 0.0
    for k in picks:
        o = P.pconv(cvs, pts)
        o.push_ir(irs[k])
        want = np.stack([o.convolution(x[t, k].cpu().numpy()) for t in range(2)]).astype(np.float64)
        got = y[:, k].cpu().numpy().astype(np.float64)
        worst = max(worst, float(np.linalg.norm(got - want) / np.linalg.norm(want)))
    return {"rel_l2_vs_oracle": worst, "channels_checked": picks, "blocks": 2, "tolerance": 1e-5}


def bench_pconv_single_process(eng, devices, steps, warmup, channels_per_gpu=CHANNELS_PER_GPU, cvs=CVS, pts=PTS):
    """The weak workload through ONE multi-GPU handle in ONE process: host buffers in, host buffers out."""
    import numpy as np
    import torch

    n = len(devices)
    channels = channels_per_gpu * n
    conv = eng.Clpconv(0, cvs, pts, channels=channels, devices=devices)
    if conv.get_cl_err():
        raise RuntimeError("Clpconv(devices=...): " + eng.cl_error_string(conv.get_cl_err()))
    # decaying-noise IRs generated on the host in float32: one GPU's worth (1.9 GB), uploaded to every device's shard
    rng = np.random.default_rng(7000)
    env = np.exp(-6.9078 * np.arange(cvs, dtype=np.float32) / cvs)
    ir = np.empty((channels_per_gpu, cvs), np.float32)
    for c0 in range(0, channels_per_gpu, 64):
        blk = rng.standard_normal((min(64, channels_per_gpu - c0), cvs), dtype=np.float32) * env
        ir[c0:c0 + blk.shape[0]] = blk / np.linalg.norm(blk, axis=1, keepdims=True)
    for g in range(n):
        assert conv.push_ir_shard(g, ir) == 0
    del ir
    hx = torch.empty(4, channels, pts).pin_memory()
    hx.uniform_(-1, 1)
    hy = torch.empty(channels, pts).pin_memory()
    hx_np, hy_np = hx.numpy(), hy.numpy()
    for i in range(warmup):
        assert conv.convolution(hy_np, hx_np[i % 4]) == 0
    t0 = time.perf_counter()
    for i in range(steps):
        assert conv.convolution(hy_np, hx_np[i % 4]) == 0
    ms = (time.perf_counter() - t0) * 1e3 / steps
    conv.close()
    return {"value": channels * (pts / SR) / (ms * 1e-3), "unit": "realtime_channels_48k", "ms_per_step": ms,
            "devices": n, "channels_total": channels, "h2d_bytes_per_step": channels * pts * 4,
            "d2h_bytes_per_step": channels * pts * 4}


def bench_rfft(eng, local, rank, world, steps, warmup, size=RFFT_SIZE, batch=RFFT_BATCH, e2e=False):
    import torch

    from opencl_fft_b200.shard import max_over_ranks

    plan = eng.Clrfft(local, size, True, max_batch=batch)
    if plan.get_error():
        raise RuntimeError("Clrfft: " + eng.cl_error_string(plan.get_error()))
    g = torch.Generator(device="cuda").manual_seed(6000 + rank)
    nbuf = 3  # 3 x 256 MiB in + 3 x 256 MiB out: successive steps never find their data in the 126 MB L2
    x = (torch.rand(nbuf, batch, size, generator=g, device="cuda") * 2 - 1).contiguous()
    y = torch.empty(nbuf, batch, size, device="cuda")

    def step(i):
        assert plan.transform_dev(x[i % nbuf], y[i % nbuf], batch) == 0

    with ClockSampler(local) as clk:
        ms = event_time_ms(step, steps, warmup)
    ms_max = max_over_ranks(ms)
    bytes_per_step = batch * 8 * size  # 4*size in + 8*(size/2) out
    res = {"ms_per_step": ms_max / steps, "bytes_per_step": bytes_per_step, "clocks": clk.summary(),
           "batch": batch, "size": size}
    if e2e:
        hx = torch.empty(batch, size).pin_memory()
        hx.copy_(x[0].cpu())
        hc = torch.empty(batch, size).pin_memory()
        a, c = hx.numpy().reshape(-1), hc.numpy().reshape(-1).view("complex64")
        for _ in range(2):
            assert plan.transform(c, a) == 0
        barrier()
        n = max(3, min(steps, 10))
        t0 = time.perf_counter()
        for _ in range(n):
            assert plan.transform(c, a) == 0
        t1 = time.perf_counter()
        res["e2e_ms_per_step"] = max_over_ranks((t1 - t0) * 1e3) / n
        res["h2d_bytes_per_step"] = batch * size * 4
        res["d2h_bytes_per_step"] = batch * size * 4
    plan.close()
    return res


def bench_cfft1024(eng, local, steps, warmup):
    import torch

    N, batch = 1024, 65536  # 512 MiB in, 512 MiB out
    plan = eng.Clcfft(local, N, True, max_batch=1)
    x = torch.randn(2, batch, N, 2, device="cuda")
    y = torch.empty_like(x)

    def step(i):
        assert plan.transform_dev(x[i % 2], y[i % 2], batch) == 0

    ms = event_time_ms(step, steps, warmup)
    plan.close()
    return {"ms_per_step": ms / steps, "bytes_per_step": batch * 16 * N}


def bench_dconv(eng, local, steps, warmup):
    import torch

    irsize, vsize, ch, nblocks = 4096, 256, 64, 375  # config 4, 2 s of audio per launch
    conv = eng.Cldconv(local, irsize, vsize, channels=ch, max_blocks=1)
    ir = torch.randn(ch, irsize, device="cuda") / 64
    assert conv.push_ir_dev(ir, irsize) == 0
    x = torch.rand(ch, nblocks * vsize, device="cuda") * 2 - 1
    y = torch.empty_like(x)

    def step(i):
        assert conv.convolution_dev(y, x, nblocks=nblocks) == 0

    ms = event_time_ms(step, steps, warmup)
    # single-block latency (the streaming call of the reference): one 256-sample block, 64 channels
    xb = x[:, :vsize].contiguous()
    yb = torch.empty_like(xb)

    def step1(i):
        assert conv.convolution_dev(yb, xb, nblocks=1) == 0

    ms1 = event_time_ms(step1, 50, 5)
    conv.close()
    flop = 2.0 * irsize * vsize * nblocks * ch
    return {"ms_per_step": ms / steps, "flop_per_step": flop, "single_block_us": ms1 / 50 * 1e3}


def bench_rfft4096(eng, local, steps, warmup):
    """BASELINE config 2 shape, batched: 32768 x 4096-point real FFT, forward then inverse (round trip)."""
    import torch

    size, batch = 4096, 32768  # 512 MiB per direction
    f = eng.Clrfft(local, size, True, max_batch=1)
    i = eng.Clrfft(local, size, False, max_batch=1)
    x = torch.rand(2, batch, size, device="cuda") * 2 - 1
    y = torch.empty_like(x)

    def step(k):
        assert f.transform_dev(x[k % 2], y[k % 2], batch) == 0
        assert i.transform_dev(y[k % 2], y[k % 2], batch) == 0

    ms = event_time_ms(step, steps, warmup)
    f.close()
    i.close()
    return {"ms_per_step": ms / steps, "bytes_per_step": 2 * batch * 8 * size}


def bench_latency(eng, local):
    """Per-call latency of the reference-facing synchronous host API at the reference's own shapes
    (batch 1 / mono): what a Csound performance thread would see per call."""
    import numpy as np

    out = {}
    rng = np.random.default_rng(0)

    def timeit(fn, n=300):
        for _ in range(20):
            fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n * 1e6

    c = eng.Clcfft(local, 1024, True)
    x = (rng.uniform(-1, 1, 1024) + 1j * rng.uniform(-1, 1, 1024)).astype(np.complex64)
    out["cfft1024_batch1_us"] = timeit(lambda: c.transform(x))
    c.close()
    f, i = eng.Clrfft(local, 4096, True), eng.Clrfft(local, 4096, False)
    r = rng.uniform(-1, 1, 4096).astype(np.float32).view(np.complex64)
    out["rfft4096_roundtrip_batch1_us"] = timeit(lambda: (f.transform(r), i.transform(r)))
    f.close()
    i.close()
    p = eng.Clpconv(local, 96000, 512)
    p.push_ir((rng.standard_normal(96000) * 0.01).astype(np.float32))
    xin, yout = rng.uniform(-1, 1, 512).astype(np.float32), np.zeros(512, np.float32)
    us = timeit(lambda: p.convolution(yout, xin))
    out["pconv_96000x512_mono_block_us"] = us
    out["pconv_96000x512_mono_realtime_ratio"] = (512 / SR * 1e6) / us
    p.close()
    d = eng.Cldconv(local, 4096, 256, channels=64)
    d.push_ir((rng.standard_normal((64, 4096)) / 64).astype(np.float32))
    xin, yout = rng.uniform(-1, 1, (64, 256)).astype(np.float32), np.zeros((64, 256), np.float32)
    us = timeit(lambda: d.convolution(yout, xin))
    out["dconv_4096x256x64ch_block_us"] = us
    out["dconv_4096x256x64ch_realtime_ratio"] = (256 / SR * 1e6) / us
    d.close()
    # three corners of the reference's own benchmark grid (csound/tests.py:5-36: mono, time-varying, partition M x IR
    # length L; the full grid: tools/rt_ratio_grid.py -> profiles/r02_rt_ratio_grid.txt)
    grid = {}
    for M, L in ((512, 1 << 22), (2048, 1 << 22), (8192, 1 << 20)):
        g = eng.Clpconv(local, L, M)
        a, b = rng.uniform(-1, 1, M).astype(np.float32), (rng.uniform(-1, 1, M) * 0.01).astype(np.float32)
        y = np.zeros(M, np.float32)
        us = timeit(lambda: g.convolution(y, a, b), n=100)
        grid[f"M{M}_L{L}"] = {"block_us": us, "realtime_ratio": (M / SR * 1e6) / us}
        g.close()
    out["pconv_tv_mono_grid"] = grid
    return out


def cpu_reference_latency():
    """The same four calls on the CPU reference (one thread, plan / object construction excluded)."""
    import numpy as np

    import oracle

    impl = oracle.best()
    rng = np.random.default_rng(0)
    out = {"kind": impl.kind}
    x = (rng.uniform(-1, 1, (200, 1024)) + 1j * rng.uniform(-1, 1, (200, 1024))).astype(np.complex64)
    out["cfft1024_batch1_us"] = impl.cfft_run(x, True, 1)[0] / 200 * 1e6
    r = rng.uniform(-1, 1, (100, 4096)).astype(np.float32)
    secs, spec = impl.rfft_run(r, True, 1)
    out["rfft4096_roundtrip_batch1_us"] = (secs + impl.rfft_run(spec, False, 1)[0]) / 100 * 1e6
    ir = (rng.standard_normal((1, 96000)) * 0.01).astype(np.float32)
    xs = rng.uniform(-1, 1, (1, 40 * 512)).astype(np.float32)
    out["pconv_96000x512_mono_block_us"] = impl.pconv_run(96000, 512, ir, xs, 1)[0] / 40 * 1e6
    h = (rng.standard_normal((1, 4096)) / 64).astype(np.float32)
    xd = rng.uniform(-1, 1, (1, 20 * 256)).astype(np.float32)
    # 64 channels on one thread, as a single Csound performance thread would run them
    out["dconv_4096x256x64ch_block_us"] = impl.dconv_run(4096, 256, h, xd, 1)[0] / 20 * 64 * 1e6
    return out


def cpu_baseline_pconv(threads, cvs=CVS, pts=PTS, blocks=None):
    """The reference's own implementation on the host cores: `threads` channels (one per thread)."""
    import numpy as np

    import oracle

    impl = oracle.best()
    nparts = cvs // pts
    if blocks is None:
        blocks = 1600  # ~10 s wall on 16 cores
    rng = np.random.default_rng(7000)
    n = np.arange(cvs)
    ir = (rng.standard_normal((threads, cvs)) * np.exp(-6.9078 * n / cvs)).astype(np.float32)
    x = rng.uniform(-1, 1, (threads, blocks * pts)).astype(np.float32)
    secs, _ = impl.pconv_run(cvs, pts, ir, x, threads)
    t_step = secs / blocks
    return {
        "value": threads * (pts / SR) / t_step,
        "unit": "realtime_channels_48k",
        "cores": threads,
        "kind": impl.kind,
        "sample": f"{threads} channels x {blocks} blocks of {pts} samples, {nparts} partitions each, "
                  f"{secs:.2f} s wall",
        "seconds": secs,
    }


def cpu_baseline_rfft(threads, size=RFFT_SIZE, batch=None, reps=24):
    """`reps` passes over a batch of 128 x threads transforms (512 MB of input at 16 threads): ~10 s of CPU work"""
    import numpy as np

    import oracle

    impl = oracle.best()
    if batch is None:
        batch = 128 * threads
    rng = np.random.default_rng(6000)
    x = rng.uniform(-1, 1, (batch, size)).astype(np.float32)
    secs = 0.0
    for _ in range(reps):
        secs += impl.rfft_run(x, True, threads)[0]
    return {"value": reps * batch * 8 * size / secs / 1e9, "unit": "GB/s", "cores": threads, "kind": impl.kind,
            "sample": f"{reps} x {batch} transforms of {size} real points, {secs:.2f} s wall", "seconds": secs}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation, same metric/unit/config as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    fn = cpu_baseline_pconv if args.workload == "pconv" else cpu_baseline_rfft
    small = {"blocks": 160} if args.workload == "pconv" else {"batch": 32 * threads, "reps": 4}
    for _ in range(args.warmup):
        fn(threads, **small)
    vals, secs = [], 0.0
    for _ in range(args.steps):
        r = fn(threads, **small)
        vals.append(r["value"])
        secs += r["seconds"]
    v = statistics.median(vals)
    line = {
        "impl": "reference",
        "metric": "pconv_realtime_channels_48k" if args.workload == "pconv" else "batched_rfft_GBps",
        "value": v, "unit": r["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload),
        "cpu_baseline": {"value": v, "unit": r["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": v, "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(workload):
    if workload == "pconv":
        return {"workload": "BASELINE configs[4] (partitioned-conv half): 1024 channels per GPU x 480000-tap (10 s) IRs, "
                            "512-sample partitions (937), one 512-sample block per channel per step",
                "channels_per_gpu": CHANNELS_PER_GPU, "ir_taps": CVS, "partition": PTS, "sample_rate": 48000,
                "l2_policy": "working set 7.9 GB per step >> 126 MB L2, no flush needed",
                "parallelism": "channels sharded across GPUs, no collective"}
    return {"workload": "BASELINE configs[4] (FFT half): 1024 x 65536-point real FFT per GPU, forward",
            "batch_per_gpu": RFFT_BATCH, "size": RFFT_SIZE,
            "l2_policy": "3 rotating input/output buffer pairs of 256 MiB each, larger than the 126 MB L2",
            "parallelism": "transform batch sharded across GPUs, no collective"}


def fft_objects(eng, local, rank, world, steps, warmup, peak, peak_src, batch, e2e):
    """The FFT half of the metric: the line's `fft` object (value over all GPUs, roofline of this rank's launch)."""
    f = bench_rfft(eng, local, rank, world, steps, warmup, batch=batch, e2e=e2e)
    t = f["ms_per_step"] * 1e-3
    achieved = f["bytes_per_step"] / t / 1e9
    obj = {"metric": "batched_rfft_GBps", "value": f["bytes_per_step"] * world / t / 1e9, "unit": "GB/s",
           "ms_per_step": f["ms_per_step"], "batch_per_gpu": batch, "size": RFFT_SIZE, "clocks": f["clocks"],
           "gpu_launches": steps if batch >= 96 else 2 * steps,
           "roofline": {"bound": "hbm", "kernel": RFFT_KERNEL if batch >= 96 else RFFT_KERNEL_SMALL, "achieved": achieved,
                        "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": ncu_traffic("rfft65536_bytes_per_step") if batch == RFFT_BATCH else None,
                        "algorithmic_bytes_per_launch": f["bytes_per_step"], "peak_source": peak_src}}
    if e2e:
        obj["e2e"] = {"value": f["bytes_per_step"] * world / (f["e2e_ms_per_step"] * 1e-3) / 1e9, "unit": "GB/s",
                      "h2d_bytes_per_step": f["h2d_bytes_per_step"], "d2h_bytes_per_step": f["d2h_bytes_per_step"],
                      "ms_per_step": f["e2e_ms_per_step"]}
    return obj


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pconv", choices=["pconv", "rfft"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="which split the headline value is: 1024 channels per GPU (weak) or 1024 in total (strong); "
                         "the other one is reported beside it")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single-process", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch

    import opencl_fft_b200 as eng  # raises if the CUDA library is not built: no fallback

    rank, world, local = dist_setup(args.gpus)
    peak, peak_src, sm_max = measured_peaks()
    k = max(5, min(args.steps, 20))  # steps of the measurements that are not the headline
    first = rank == 0 and world == 1

    # ---- partitioned convolution: weak (1024 channels per GPU) and the literal configs[4] split (1024 in total)
    def pconv_obj(channels, steps, e2e, spot):
        r = bench_pconv(eng, local, rank, world, steps, args.warmup, channels=channels, e2e=e2e, spot_check=spot)
        t = r["ms_per_step"] * 1e-3
        achieved = r["bytes_per_launch"] / (r["ms_per_step_rank"] * 1e-3) / 1e9
        o = {"metric": "pconv_realtime_channels_48k", "value": channels * world * (PTS / SR) / t,
             "unit": "realtime_channels_48k", "ms_per_step": r["ms_per_step"], "channels_per_gpu": channels,
             "channels_total": channels * world, "clocks": r["clocks"], "gpu_launches": r["launches"],
             "roofline": {"bound": "hbm", "kernel": "pconv_step_kernel<9,false>", "achieved": achieved, "peak": peak,
                          "unit": "GB/s", "frac": achieved / peak,
                          "traffic": ncu_traffic("pconv_step_bytes_per_launch") if channels == CHANNELS_PER_GPU else None,
                          "algorithmic_bytes_per_launch": r["bytes_per_launch"], "peak_source": peak_src}}
        if e2e:
            o["e2e"] = {"value": channels * world * (PTS / SR) / (r["e2e_ms_per_step"] * 1e-3), "unit": o["unit"],
                        "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": r["d2h_bytes_per_step"],
                        "ms_per_step": r["e2e_ms_per_step"], "host_buffers": "pinned",
                        "same_bits_as_device_path": r["e2e_matches_device_path"],
                        "pageable_host_buffers": {"value": channels * world * (PTS / SR) / (r["e2e_pageable_ms_per_step"] * 1e-3),
                                                  "ms_per_step": r["e2e_pageable_ms_per_step"]}}
        if "parity_spot_check" in r:
            o["parity_spot_check"] = r["parity_spot_check"]
        return o

    head_p = args.workload == "pconv"
    strong_ch, strong_b = CHANNELS_PER_GPU // world, RFFT_BATCH // world
    weak_p = pconv_obj(CHANNELS_PER_GPU, args.steps if head_p else k, True, first and not args.no_cpu_baseline)
    strong_p = pconv_obj(strong_ch, k, False, False) if world > 1 else None
    weak_f = fft_objects(eng, local, rank, world, args.steps if not head_p else k, 3, peak, peak_src, RFFT_BATCH, True)
    strong_f = fft_objects(eng, local, rank, world, k, 3, peak, peak_src, strong_b, False) if world > 1 else None

    def strong_view(o, weak):
        if o is None:  # one GPU: the two splits coincide
            o = weak
        return {key: o[key] for key in ("metric", "value", "unit", "ms_per_step") if key in o} | {
            "per_gpu": o.get("channels_per_gpu", o.get("batch_per_gpu")), "roofline_frac": o["roofline"]["frac"],
            "kernel": o["roofline"]["kernel"]}

    strong = {"note": "BASELINE configs[4] literally: 1024 channels / 1024 transforms in TOTAL, split evenly over the GPUs",
              "pconv": strong_view(strong_p, weak_p), "rfft": strong_view(strong_f, weak_f)}

    # ---- one process, all GPUs, through the library's multi-GPU handle (rank 0 alone; the others wait)
    single = None
    barrier()
    if rank == 0 and not args.no_single_process and head_p:
        try:
            single = bench_pconv_single_process(eng, list(range(world)), max(5, min(args.steps, 50)), 3)
        except Exception as e:  # reported, never hidden
            single = {"error": str(e)}
    cpu_barrier()
    barrier()

    secondary = {}
    if not args.no_secondary and first:
        c = bench_cfft1024(eng, local, k, 3)
        gbs = c["bytes_per_step"] / (c["ms_per_step"] * 1e-3) / 1e9
        secondary["batched_cfft_1024x65536"] = {"value": gbs, "unit": "GB/s", "ms_per_step": c["ms_per_step"],
                                                "roofline_frac": gbs / peak}
        q = bench_rfft4096(eng, local, k, 3)
        gbs = q["bytes_per_step"] / (q["ms_per_step"] * 1e-3) / 1e9
        secondary["batched_rfft_4096x32768_roundtrip"] = {"value": gbs, "unit": "GB/s", "ms_per_step": q["ms_per_step"],
                                                          "roofline_frac": gbs / peak}
        secondary["host_api_latency"] = bench_latency(eng, local)
        if not args.no_cpu_baseline:
            secondary["host_api_latency"]["cpu_reference_1_thread"] = cpu_reference_latency()
        d = bench_dconv(eng, local, k, 3)
        tf = d["flop_per_step"] / (d["ms_per_step"] * 1e-3) / 1e12
        fp32_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        fp32_peak, fp32_src = fp32_nominal, "nominal"
        pk = os.path.join(ROOT, "profiles", "fp32_peak.json")
        if os.path.exists(pk):
            fp32_peak, fp32_src = float(json.load(open(pk))["fp32_fma_tflops"]), "measured (tools/fma_peak.cu, profiles/fp32_peak.json)"
        secondary["dconv_4096x256x64ch"] = {"value": tf, "unit": "TFLOP/s fp32", "ms_per_step": d["ms_per_step"],
                                            "roofline": {"bound": "fp32_fma", "achieved": tf, "peak": fp32_peak,
                                                         "unit": "TFLOP/s", "frac": tf / fp32_peak, "peak_source": fp32_src},
                                            "fp32_peak_nominal": fp32_nominal,
                                            "realtime_channels_48k": 64 * (375 * 256 / SR) / (d["ms_per_step"] * 1e-3),
                                            "single_block_latency_us": d["single_block_us"]}

    cpu = cpu_f = None
    if not args.no_cpu_baseline and first:
        threads = os.cpu_count() or 1
        cpu = cpu_baseline_pconv(threads)
        cpu.pop("seconds", None)
        cpu_f = cpu_baseline_rfft(threads)
        cpu_f.pop("seconds", None)
        weak_f["cpu_baseline"] = cpu_f

    if rank == 0:
        strong_head = args.scaling == "strong" and world > 1
        head = (strong_p if strong_head else weak_p) if head_p else (strong_f if strong_head else weak_f)
        other = weak_f if head_p else weak_p
        cfg = workload_config(args.workload)
        if strong_head:
            cfg["workload"] += f" -- STRONG split: {head.get('channels_per_gpu', head.get('batch_per_gpu'))} per GPU"
        line = {
            "metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if strong_head else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "clocks": head["clocks"],
            "e2e": head.get("e2e") or (weak_p if head_p else weak_f)["e2e"],
            "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
        }
        if "parity_spot_check" in weak_p:
            line["parity_spot_check"] = weak_p["parity_spot_check"]
        if head_p:
            line["fft"] = {kk: v for kk, v in other.items() if kk != "metric"} | {"metric": other["metric"]}
            if cpu is not None:
                line["cpu_baseline"] = cpu
        else:
            line["pconv"] = other
            if cpu_f is not None:
                line["cpu_baseline"] = cpu_f
        line["strong"] = strong
        if single is not None:
            line["e2e_single_process"] = single
        if secondary:
            line["secondary"] = secondary
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
