#!/usr/bin/env python
"""Burst-mode A/B of the dispatch choices that have run-time hooks (radix32_mask, pipe_mask, wpipe): for the sizes where more
than one kernel exists, time every alternative on the same buffers with an idle pause before each cell and the best of three
timed groups, so that no cell is measured under the board's power cap (the back-to-back sweeps of round 1 were: see
profiles/r02_power_cap.txt).  Prints GB/s per alternative and the winner.  GPU only."""
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import chowdsp_fft_b200 as cf  # noqa: E402

total_floats = 2 * 2**30 // 4
x = torch.rand(total_floats, device="cuda") * 2 - 1
y = torch.empty_like(x)
stream = torch.cuda.current_stream()


def measure(plan, batch, nfl, direction, ordered, steps=20, groups=3, pause=1.0):
    torch.cuda.synchronize()
    time.sleep(pause)
    for _ in range(3):
        cf.fft_transform_batched(plan, x, y, batch, nfl, nfl, direction, ordered, stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = float("inf")
    for _ in range(groups):
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            cf.fft_transform_batched(plan, x, y, batch, nfl, nfl, direction, ordered, stream)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    return batch * nfl * 8 / (best * 1e-3) / 1e9, cf.last_kernel()


ALTS = {  # name -> tuning keys
    "r16 nopipe": dict(radix32_mask=0, pipe_mask=0, wpipe=2),
    "r32 nopipe": dict(radix32_mask=0x7FFFFFFF, pipe_mask=0, wpipe=2),
    "pipe": dict(radix32_mask=0x7FFFFFFF, pipe_mask=0xFFFF, wpipe=2),
    "wpipe": dict(radix32_mask=0x7FFFFFFF, pipe_mask=0, wpipe=3),
    "default": dict(radix32_mask=-1, pipe_mask=-1, wpipe=-1),
}
for is_c, N in [(True, 512), (True, 1024), (True, 8192), (True, 16384), (False, 1024), (False, 2048), (False, 16384), (False, 32768)]:
    nfl = 2 * N if is_c else N
    plan = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL, True)
    batch = total_floats // nfl
    for direction in (cf.FFT_FORWARD, cf.FFT_BACKWARD):
        for ordered in (True, False):
            res, seen = {}, {}
            for name, keys in ALTS.items():
                for k, v in keys.items():
                    cf.set_tuning(k, v)
                g, kern = measure(plan, batch, nfl, direction, ordered)
                if kern in seen and name != "default":
                    continue  # this alternative routes to a kernel already timed
                seen[kern] = name
                res[name] = (g, kern)
            kind = "C2C" if is_c else ("R2C" if direction == 0 else "C2R")
            best = max((n for n in res if n != "default"), key=lambda n: res[n][0])
            line = "  ".join(f"{n}={res[n][0]:6.0f}" for n in res)
            print(f"{kind} N={N:6d} {'fwd' if direction == 0 else 'bwd'} {'ordered' if ordered else 'w8     '}  {line}  best: {best} ({res[best][1]})  default: {res['default'][1]}", flush=True)
    cf.fft_destroy_setup(plan)
