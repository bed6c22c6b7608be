"""Distributed four-step FFT: ONE complex transform of N = 2^n points (n >= 21) over G GPUs, one process
per GPU.  The local phases are C-ABI calls (fft_dist_phase).  The exchange between phase 0 and phase 1 is
either FUSED into phase 0 (exchange="peer", default: the kernel stores every row block straight into its owner's
buffer through NVLink peer memory, so the transfer overlaps the butterflies; torch.distributed only carries the
IPC handles at set-up and a one-element all-reduce as the barrier) or an NCCL all_to_all_single
(exchange="nccl", the baseline).  See include/chowdsp_fft_b200.h for the data contracts.

    input  (per rank)  column block  A[n1][c] = x[n1*S1 + rank*S1/G + c],  shape [L1, S1/G] complex
    output (per rank)  transposed-out  out[q][k] = X[(rank*L1/G + k) + L1*q], shape [S1, L1/G] complex
                       natural() redistributes it into contiguous blocks of X with a second all-to-all
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
        dev = torch.device("cuda", torch.cuda.current_device())
        self.nat = torch.empty(self.rows * self.S1 * 2, device=dev)         # natural rows [rows][S1]
        self.step = 0
        if exchange == "nccl":
            self.send = torch.empty(self.L1 * self.cols * 2, device=dev)   # phase-0 output / all-to-all source
            self.recv = torch.empty_like(self.send)                         # exchange layout [G][rows][cols]
            return
        # peer exchange: two receive buffers (step parity) so that one barrier per transform is enough; every rank
        # maps every other rank's buffers through CUDA IPC
        nbytes = self.L1 * self.cols * 8
        self.own = [api.fft_dist_alloc(nbytes) for _ in range(2)]
        self.mapped: list[int] = []
        if world > 1:
            handles: list = [None] * world
            dist.all_gather_object(handles, [api.fft_dist_ipc_export(p) for p in self.own], group=group)
            self.peers = []
            for b in range(2):
                row = []
                for h in range(world):
                    if h == rank:
                        row.append(self.own[b])
                    else:
                        row.append(api.fft_dist_ipc_open(handles[h][b]))
                        self.mapped.append(row[-1])
                self.peers.append(row)
            self.token = torch.zeros(1, device=dev)
        else:
            self.peers = [[self.own[0]], [self.own[1]]]

    @property
    def local_floats(self) -> int:
        return 2 * self.L1 * self.cols

    def exchange_bytes(self) -> int:
        """bytes this rank sends over NVLink in one all-to-all"""
        return (self.world - 1) * self.rows * self.cols * 8

    def forward(self, x_cols: torch.Tensor, out_t: torch.Tensor, direction: int = api.FFT_FORWARD, stream=None):
        """x_cols: [L1, S1/G] complex as float32 pairs (flat ok); out_t: [S1, L1/G] complex (flat ok)."""
        st = stream or torch.cuda.current_stream()
        # the collectives below order themselves against torch's CURRENT stream, so the phase kernels must run on it too:
        # make `st` current for the whole body (a caller-supplied stream other than the current one would otherwise let
        # phase 1 read buffers that peers are still writing)
        with torch.cuda.stream(st):
            return self._forward_on(st, x_cols, out_t, direction)

    def _forward_on(self, st, x_cols: torch.Tensor, out_t: torch.Tensor, direction: int):
        if self.exchange == "peer":
            b = self.step & 1
            self.step += 1
            api.fft_dist_phase0_peer(self.plan, self.rank, self.world, x_cols, self.peers[b], direction, st)
            if self.world > 1:
                dist.all_reduce(self.token, group=self.group)  # stream-ordered barrier: every rank's stores have landed
            src = self.own[b]
        else:
            api.fft_dist_phase(self.plan, 0, self.rank, self.world, x_cols, self.send, direction, st)
            if self.world > 1:
                dist.all_to_all_single(self.recv, self.send, group=self.group)
                src = self.recv
            else:
                src = self.send
        api.fft_dist_phase(self.plan, 1, self.rank, self.world, src, self.nat, direction, st)
        api.fft_dist_phase(self.plan, 2, self.rank, self.world, self.nat, out_t, direction, st)
        return out_t

    def natural(self, out_t: torch.Tensor) -> torch.Tensor:
        """Second all-to-all: transposed-out -> this rank's contiguous block X[rank*N/G : (rank+1)*N/G]."""
        G = self.world
        t = out_t.view(self.S1, self.rows, 2)
        if G == 1:
            return t.reshape(-1)
        got = torch.empty_like(t)
        dist.all_to_all_single(got.view(-1), t.reshape(-1), group=self.group)  # chunk h = rows q of block `rank`, k of rank h
        # got[h][q_local][k_local] -> X[(h*rows + k) + L1*(rank*S1/G + q_local)]
        return got.view(G, self.cols, self.rows, 2).permute(1, 0, 2, 3).reshape(-1)

    def close(self):
        if self.exchange == "peer":
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)  # nobody unmaps or frees while a peer may still be storing
            for p in self.mapped:
                api.fft_dist_ipc_close(p)
            for p in self.own:
                api.fft_dist_free(p)
            self.mapped, self.own = [], []
        api.fft_destroy_setup(self.plan)


def column_block(x_full: torch.Tensor, L1: int, S1: int, rank: int, world: int) -> torch.Tensor:
    """Helper for tests: the column block of a full natural-order signal (complex as [..., 2] float32)."""
    cols = S1 // world
    return x_full.view(L1, S1, 2)[:, rank * cols:(rank + 1) * cols].contiguous()
