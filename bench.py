#!/usr/bin/env python
"""Benchmark of the chowdsp_fft hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload =
BASELINE.json configs[1]: batched complex C2C N=4096 x 65536 transforms fp32, ordered, 1 GPU
(4 GiB algorithmic bytes per step, working set far larger than the 126 MB L2, so no L2 flush is
needed between iterations).  Prints ONE JSON line on rank 0.

  value        whole-job throughput, inputs resident in HBM, CUDA-event timed on the launch stream
  e2e          same metric through the C-ABI with HOST (pinned) buffers: H2D + transform + D2H timed
  roofline     algorithmic bytes / launch duration vs the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline the reference's own AVX build (oracle/_ref) one thread per host core, bounded sample
--impl reference runs ONLY that CPU arm (rank 0), as the driver's comparison line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback

# name -> (N, is_complex, batch per GPU, ordered, description)
WORKLOADS = {
    "c2c4096": dict(N=4096, is_complex=True, batch=65536, ordered=True,
                    desc="batched C2C N=4096 x 65536 fp32, ordered, forward (BASELINE configs[1])"),
    "c2c4096_unordered": dict(N=4096, is_complex=True, batch=65536, ordered=False,
                              desc="batched C2C N=4096 x 65536 fp32, unordered (reference W=8 layout), forward"),
    "c2c1024": dict(N=1024, is_complex=True, batch=262144, ordered=True, desc="batched C2C N=1024 x 262144"),
    "c2c16384": dict(N=16384, is_complex=True, batch=16384, ordered=True, desc="batched C2C N=16384 x 16384"),
    "r2c2048": dict(N=2048, is_complex=False, batch=524288, ordered=True, desc="batched R2C N=2048 x 524288"),
    "r2c8192": dict(N=8192, is_complex=False, batch=131072, ordered=False, desc="batched R2C N=8192 x 131072 unordered"),
}


def algorithmic_bytes(N: int, is_complex: bool) -> int:
    """C2C 16 N bytes per transform (8N in + 8N out); R2C/C2R 8 N bytes."""
    return 16 * N if is_complex else 8 * N


def flops(N: int, is_complex: bool) -> float:
    return (5.0 if is_complex else 2.5) * N * math.log2(N)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML during the timed region."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _run(self):
        n = self._nvml
        names = {
            n.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            n.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            n.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            n.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            n.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_arm(wl, seconds_target: float, steps: int = 1, warmup: int = 0):
    """Times the unmodified reference (oracle/_ref) with one thread per host core on a bounded sample
    of the workload.  Returns (GB/s, cores, sample description, ms per step, SIMD width)."""
    from oracle import oracle as o

    ref = o.load_ref()
    if ref is None:
        raise RuntimeError("oracle/_ref/libchowdsp_fft_ref.so is missing (build it with make -C oracle)")
    N, is_c, ordered = wl["N"], wl["is_complex"], wl["ordered"]
    nfl = 2 * N if is_c else N
    cores = ref.hardware_threads()
    try:
        cores = min(cores, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    # working set: enough transforms that every core streams from its own slice (>= 256 per core)
    sample = min(wl["batch"], max(cores * 256, 2048))
    rng = np.random.default_rng(42)
    xin = o.aligned_copy(rng.uniform(-1, 1, sample * nfl).astype(np.float32))
    out = o.aligned_empty(sample * nfl)
    t1 = ref.transform_timed(xin, out, N, is_c, False, ordered, sample, nfl, nfl, cores)  # also warms up
    reps = max(1, int(seconds_target / max(t1, 1e-6) / max(1, steps + warmup)))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for _ in range(reps):
            ref.transform_timed(xin, out, N, is_c, False, ordered, sample, nfl, nfl, cores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    per_step = float(np.mean(times))
    gbs = sample * reps * algorithmic_bytes(N, is_c) / per_step / 1e9
    width = o.simd_width(N, is_c, True) * 4
    desc = f"{sample} transforms x {reps} passes per step, {cores} threads (one per core, pinned), reference AVX build W={width}B"
    return gbs, cores, desc, per_step * 1e3, width


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2c4096", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]
    N, is_c, batch, ordered = wl["N"], wl["is_complex"], wl["batch"], wl["ordered"]
    nfl = 2 * N if is_c else N
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric = "batched fp32 FFT throughput, algorithmic bytes (in+out) per second"
    config = {"workload": wl["desc"], "N": N, "transform": "C2C" if is_c else "R2C", "batch_per_gpu": batch,
              "ordered": ordered, "l2_policy": "inputs+outputs (%.1f GiB per GPU) far exceed the 126 MB L2; no flush needed" % (2 * batch * nfl * 4 / 2**30 / (1 if is_c else 1)),
              "sharding": "independent transforms split by batch across GPUs, no collectives"}

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        gbs, cores, desc, ms, width = cpu_reference_arm(wl, seconds_target=20.0, steps=max(1, args.steps), warmup=args.warmup)
        line = {"impl": "reference", "metric": metric, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "gflops": gbs / algorithmic_bytes(N, is_c) * flops(N, is_c),
                "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": desc},
                "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist

    import chowdsp_fft_b200 as cf

    if not torch.cuda.is_available() or not cf.device_available():
        raise SystemExit("bench.py: no CUDA device; chowdsp_fft_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    setup = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL, True)
    gen = torch.Generator(device="cuda").manual_seed(42 + rank)
    x = torch.rand(batch, nfl, device="cuda", generator=gen) * 2 - 1
    y = torch.empty_like(x)
    stream = torch.cuda.current_stream()

    def step():
        cf.fft_transform_batched(setup, x, y, batch, nfl, nfl, cf.FFT_FORWARD, ordered, stream)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = cf.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
    launches = cf.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    bytes_step = batch * algorithmic_bytes(N, is_c)
    value = world * bytes_step / (ms_step * 1e-3) / 1e9
    per_gpu = bytes_step / (ms_total / args.steps * 1e-3) / 1e9  # this rank's kernel: 1 launch per step
    peak, peak_src = hbm_peak()

    # quick parity gate beside the timing: 64 transforms of the timed output vs the oracle (rank 0)
    parity = None
    if rank == 0:
        try:
            from oracle import oracle as o

            sel = torch.arange(0, batch, max(1, batch // 64), device="cuda")[:64]
            want = o.np_transform(x[sel].cpu().numpy(), N, is_c, 8, False, ordered)
            parity = {"rel_l2_vs_oracle": o.rel_l2(y[sel].cpu().numpy(), want), "tolerance": o.parity_tol(N), "transforms": int(sel.numel())}
        except Exception as e:  # the bench number stands on its own; tests/ are the parity gate
            parity = {"error": repr(e)}

    # ---- e2e: HOST pinned buffers through the C ABI, H2D + kernel + D2H inside the timed region --------
    e2e = None
    if not args.no_e2e:
        del y
        hin, hout = cf.aligned_array(batch * nfl), cf.aligned_array(batch * nfl)
        hin.reshape(batch, nfl)[:] = x.cpu().numpy()
        k_e2e = max(2, min(5, args.steps))
        cf.fft_transform_batched(setup, hin, hout, batch, nfl, nfl, cf.FFT_FORWARD, ordered)  # warm-up (staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            cf.fft_transform_batched(setup, hin, hout, batch, nfl, nfl, cf.FFT_FORWARD, ordered)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / k_e2e], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * bytes_step / float(dt.item()) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": batch * nfl * 4, "d2h_bytes_per_step": batch * nfl * 4,
               "ms_per_step": float(dt.item()) * 1e3, "steps": k_e2e,
               "path": "fft_transform_batched(host pinned in/out): 32 MiB chunks, H2D/kernel/D2H overlapped on two streams"}
        cf.aligned_free(hin.ctypes.data)
        cf.aligned_free(hout.ctypes.data)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            gbs, cores, desc, _, _ = cpu_reference_arm(wl, seconds_target=12.0)
            cpu = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": desc}
        except Exception as e:
            cpu = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e!r}"}

    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(args.workload)
    except Exception:
        pass

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "gflops": value / algorithmic_bytes(N, is_c) * flops(N, is_c),
                "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": "GB/s", "frac": per_gpu / peak,
                             "traffic": traffic, "peak_source": peak_src, "frac_of_nominal_8TBs": per_gpu / 8000.0,
                             "kernel": "cfb::fft_kernel<%d,16,%s,%s>" % (int(math.log2(N)) - (0 if is_c else 1), "C2C_FWD" if is_c else "R2C", "false" if ordered else "true"),
                             "algorithmic_bytes_per_launch": bytes_step},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary(),
                "parity": parity}
        print(json.dumps(line))
    cf.fft_destroy_setup(setup)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
