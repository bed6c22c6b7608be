"""Distributed four-step FFT: ONE complex transform of N = 2^n points (n >= 21) over G GPUs, one process per GPU.

exchange="peer" (default) is a BINDING of the C ABI's fft_dist_* context (include/chowdsp_fft_b200.h, csrc/capi.cu):
the all-to-all is done by the phase-0 kernel's stores into peer memory over NVLink, the barrier across ranks is a flag
exchange executed in the stream, phases 1 + 2 run L2-chunked, and the optional natural-order output fuses the second
all-to-all into the last pass's peer stores.  torch.distributed only carries the opaque IPC blobs once at set-up
(any transport would do); no collective runs on the data path.

exchange="nccl" is the baseline the fused path is measured against: the three local phases as separate C-ABI calls
(fft_dist_phase) around torch.distributed.all_to_all_single.

    input  (per rank)  column block  A[n1][c] = x[n1*S1 + rank*S1/G + c],  shape [L1, S1/G] complex
    output (per rank)  transposed-out  out[q][k] = X[(rank*L1/G + k) + L1*q], shape [S1, L1/G] complex
                       natural order: this rank's contiguous block X[rank*N/G : (rank+1)*N/G]
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import api


class DistributedFFT:
    def __init__(self, n: int, rank: int, world: int, group=None, exchange: str = "peer"):
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.n, self.rank, self.world, self.group, self.exchange = n, rank, world, group, exchange
        self.N = 1 << n
        self.plan = api.fft_new_setup(self.N, api.FFT_COMPLEX, True)
        l1, l2, l3 = api.fft_large_factors(self.plan)
        if l2 == 0:
            raise api.FFTError("distributed transforms need N >= 2^21 (three-pass plan)")
        self.L1, self.L2, self.L3 = 1 << l1, 1 << l2, 1 << l3
        self.S1 = self.L2 * self.L3
        self.rows, self.cols = self.L1 // world, self.S1 // world
        self.ctx = None
        dev = torch.device("cuda", torch.cuda.current_device())
        if exchange == "nccl":
            self.nat = torch.empty(self.rows * self.S1 * 2, device=dev)    # natural rows [rows][S1]
            self.send = torch.empty(self.L1 * self.cols * 2, device=dev)   # phase-0 output / all-to-all source
            self.recv = torch.empty_like(self.send)                         # exchange layout [G][rows][cols]
            return
        self.ctx = api.fft_dist_create(self.plan, rank, world)
        if world > 1:
            blobs: list = [None] * world
            dist.all_gather_object(blobs, api.fft_dist_export(self.ctx), group=group)
            api.fft_dist_connect(self.ctx, b"".join(blobs))
            dist.barrier(group=group)  # every rank has mapped its peers before anyone stores into them
        else:
            api.fft_dist_connect(self.ctx, None)

    @property
    def local_floats(self) -> int:
        return 2 * self.L1 * self.cols

    def exchange_bytes(self) -> int:
        """bytes this rank sends over NVLink in one all-to-all"""
        return (self.world - 1) * self.rows * self.cols * 8

    def forward(self, x_cols: torch.Tensor, out_t: torch.Tensor, direction: int = api.FFT_FORWARD, stream=None, natural: bool = False, timed: bool = False):
        """x_cols: [L1, S1/G] complex as float32 pairs (flat ok); out_t: [S1, L1/G] complex (flat ok), or with natural=True
        this rank's contiguous block of X (N/G complex)."""
        st = stream or torch.cuda.current_stream()
        if self.exchange == "peer":
            api.fft_dist_transform(self.ctx, x_cols, out_t, direction, natural, timed, st)
            return out_t
        # the collective below orders itself against torch's CURRENT stream, so the phase kernels must run on it too
        with torch.cuda.stream(st):
            api.fft_dist_phase(self.plan, 0, self.rank, self.world, x_cols, self.send, direction, st)
            if self.world > 1:
                dist.all_to_all_single(self.recv, self.send, group=self.group)
                src = self.recv
            else:
                src = self.send
            api.fft_dist_phase(self.plan, 1, self.rank, self.world, src, self.nat, direction, st)
            if not natural:
                api.fft_dist_phase(self.plan, 2, self.rank, self.world, self.nat, out_t, direction, st)
                return out_t
            tmp = torch.empty(self.S1 * self.rows * 2, device=x_cols.device)
            api.fft_dist_phase(self.plan, 2, self.rank, self.world, self.nat, tmp, direction, st)
            out_t.view(-1)[:] = self.natural(tmp)
            return out_t

    def phase_ms(self) -> list[float]:
        """[phase 0 incl. peer stores, barrier wait, phases 1+2, natural-order barrier + copy] of the last timed peer transform"""
        return api.fft_dist_phase_ms(self.ctx)

    def natural(self, out_t: torch.Tensor) -> torch.Tensor:
        """Second all-to-all with NCCL: transposed-out -> this rank's contiguous block X[rank*N/G : (rank+1)*N/G]."""
        G = self.world
        t = out_t.view(self.S1, self.rows, 2)
        if G == 1:
            return t.reshape(-1)
        got = torch.empty_like(t)
        dist.all_to_all_single(got.view(-1), t.reshape(-1), group=self.group)  # chunk h = rows q of block `rank`, k of rank h
        # got[h][q_local][k_local] -> X[(h*rows + k) + L1*(rank*S1/G + q_local)]
        return got.view(G, self.cols, self.rows, 2).permute(1, 0, 2, 3).reshape(-1)

    def close(self):
        if self.ctx is not None:
            torch.cuda.synchronize()
            api.fft_dist_status(self.ctx)
            if self.world > 1:
                dist.barrier(group=self.group)  # nobody unmaps or frees while a peer may still be storing
            api.fft_dist_destroy(self.ctx)
            self.ctx = None
        api.fft_destroy_setup(self.plan)


def column_block(x_full: torch.Tensor, L1: int, S1: int, rank: int, world: int) -> torch.Tensor:
    """Helper for tests: the column block of a full natural-order signal (complex as [..., 2] float32)."""
    cols = S1 // world
    return x_full.view(L1, S1, 2)[:, rank * cols:(rank + 1) * cols].contiguous()
