#!/usr/bin/env python
"""Device-resident throughput sweep over sizes / kinds / layouts (CUDA-event timed, buffers > L2).
Prints one line per case: algorithmic GB/s and fraction of the measured HBM peak.  GPU only."""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import chowdsp_fft_b200 as cf  # noqa: E402


class GpuState:
    """SM clock / board power / throttle reasons sampled through NVML while a cell is timed (every cell of the sweep is
    judged against the HBM roofline, so a cell measured under a power cap or on a hot board has to say so)."""

    def __init__(self, index=0):
        self.clk, self.pw, self.reasons = [], [], set()
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.n = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.n = None

    def _run(self):
        n = self.n
        names = {n.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", n.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 n.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal", n.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal"}
        while not self._stop.is_set():
            try:
                self.clk.append(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                self.pw.append(n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reasons |= {v for k, v in names.items() if r & k}
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        self.t = None
        if self.n is not None:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.t is not None:
            self.t.join()

    def summary(self):
        if not self.clk:
            return dict(sm_mhz=None, power_w=None, reasons=[])
        c, p = sorted(self.clk), sorted(self.pw)
        return dict(sm_mhz=c[len(c) // 2], power_w=p[len(p) // 2] if p else None, reasons=sorted(self.reasons))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=float, default=1.0, help="GiB per buffer")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--sizes", default="")
    ap.add_argument("--json", default="")
    ap.add_argument("--radix32-mask", type=int, default=-1)
    ap.add_argument("--kinds", default="c,r")
    ap.add_argument("--layouts", default="ordered,w8,w4", help="ordered, w8 (8-lane unordered, AVX handle), w4 (4-lane unordered, SSE handle)")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE")
    ap.add_argument("--pause", type=float, default=0.0, help="idle seconds before each cell (lets a power-capped board recover)")
    ap.add_argument("--repeats", type=int, default=1, help="timed groups of --steps launches per cell; the best group is reported")
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    if args.radix32_mask >= 0:
        cf.set_tuning("radix32_mask", args.radix32_mask)
        print("radix32_mask:", args.radix32_mask)
    for kv in args.tune:
        k, v = kv.split("=")
        cf.set_tuning(k, int(v, 0))
        print("tune:", k, v)
    total_floats = int(args.bytes * 2**30 / 4)
    x = torch.rand(total_floats, device="cuda") * 2 - 1
    y = torch.empty_like(x)
    stream = torch.cuda.current_stream()
    rows = []
    sizes = [int(s) for s in args.sizes.split(",")] if args.sizes else [1 << l for l in range(5, 16)]
    layouts = args.layouts.split(",")
    for is_c in [k == "c" for k in args.kinds.split(",")]:
        for N in sizes:
            nfl = 2 * N if is_c else N
            plans = {}
            for avx in (True, False):
                try:
                    plans[avx] = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL, avx)
                except cf.FFTError:
                    pass
            if True not in plans:
                continue
            w8 = cf.fft_simd_width_bytes(plans[True]) == 32
            batch = total_floats // nfl
            for direction in (cf.FFT_FORWARD, cf.FFT_BACKWARD):
                for layout in layouts:
                    if layout == "w8" and not w8:
                        continue
                    s = plans[False] if layout == "w4" else plans[True]
                    ordered = layout == "ordered"

                    def step():
                        cf.fft_transform_batched(s, x, y, batch, nfl, nfl, direction, ordered, stream)
                    if args.pause > 0:
                        torch.cuda.synchronize()
                        time.sleep(args.pause)
                    for _ in range(3):
                        step()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ms = float("inf")
                    with GpuState() as gs:
                        for _ in range(max(1, args.repeats)):
                            torch.cuda.synchronize()
                            e0.record(stream)
                            for _ in range(args.steps):
                                step()
                            e1.record(stream)
                            torch.cuda.synchronize()
                            ms = min(ms, e0.elapsed_time(e1) / args.steps)
                    st = gs.summary()
                    gbs = batch * nfl * 8 / (ms * 1e-3) / 1e9
                    gfl = batch * (5.0 if is_c else 2.5) * N * math.log2(N) / (ms * 1e-3) / 1e9
                    row = dict(kind=("C2C" if is_c else ("R2C" if direction == 0 else "C2R")), N=N,
                               dir="fwd" if direction == 0 else "bwd", layout=layout,
                               batch=batch, ms=ms, gbs=gbs, frac=gbs / peak, frac_nominal=gbs / 8000.0, gflops=gfl, kernel=cf.last_kernel(), **st)
                    rows.append(row)
                    print(f"{row['kind']:4s} N={N:7d} {row['dir']} {row['layout']:7s} batch={batch:8d} {ms:8.4f} ms {gbs:8.1f} GB/s  frac={gbs/peak:5.3f} of measured, {gbs/8000.0:5.3f} of 8 TB/s  {gfl/1e3:6.2f} TFLOP/s  {st['sm_mhz']} MHz {st['power_w'] and round(st['power_w'])} W {','.join(st['reasons']) or '-'}  {row['kernel']}", flush=True)
            for s in plans.values():
                cf.fft_destroy_setup(s)
    if args.json:
        json.dump({"peak_gbs": peak, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
