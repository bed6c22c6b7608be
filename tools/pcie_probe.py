#!/usr/bin/env python
"""Host<->device copy ceilings of the box (pinned memory, CUDA events): H2D alone, D2H alone, both directions at once,
for several chunk sizes.  Context for bench.py's e2e line (which is PCIe-bound).  GPU only."""
import torch

total = 2 << 30
h_in = torch.empty(total, dtype=torch.uint8).pin_memory()
h_out = torch.empty(total, dtype=torch.uint8).pin_memory()
d_a = torch.empty(total, dtype=torch.uint8, device="cuda")
d_b = torch.empty(total, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(chunk, h2d, d2h):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    s1.wait_event(e0)
    s2.wait_event(e0)
    for off in range(0, total, chunk):
        if h2d:
            with torch.cuda.stream(s1):
                d_a[off:off + chunk].copy_(h_in[off:off + chunk], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[off:off + chunk].copy_(d_b[off:off + chunk], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for chunk in (8 << 20, 32 << 20, 128 << 20, 1 << 30):
    run(chunk, True, True)
    a, b, c = run(chunk, True, False), run(chunk, False, True), run(chunk, True, True)
    print(f"chunk {chunk >> 20:5d} MiB: H2D {total / a / 1e6:6.1f} GB/s   D2H {total / b / 1e6:6.1f} GB/s   both at once {total / c / 1e6:6.1f} GB/s each way "
          f"({2 * total / c / 1e6:6.1f} total)", flush=True)
