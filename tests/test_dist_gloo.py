"""Distributed four-step transform on CPU: world_size-2 gloo process group, the three local phases run from the
CUDA kernel SOURCE through the host emulator (tests/emu), the exchange between phase 0 and phase 1 over gloo
point-to-point (gloo has no all_to_all).  Checks the data contracts of include/chowdsp_fft_b200.h
(column-block input, exchange layout, transposed-out result) and the peer-store variant of phase 0, whose
stores must land exactly where the all-to-all would have put the chunks."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

fp = C.POINTER(C.c_float)
N_LOG, FACTORS = 18, (6, 6, 6)  # forced three-pass plan so that a small transform exercises the distributed path


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _emu():
    from tests.emu.build_emu import build

    L = C.CDLL(build())
    L.emu_dist_phase.argtypes = [C.c_int] * 7 + [fp, fp]
    L.emu_dist_phase0_peer.argtypes = [C.c_int] * 6 + [fp, C.POINTER(fp)]
    L.emu_dist_schedule.argtypes = [C.c_int] * 7 + [C.c_longlong, C.c_int, fp, fp, C.POINTER(fp)]
    return L


def _signal():
    rng = np.random.default_rng(7)
    return rng.uniform(-1, 1, (1 << N_LOG, 2)).astype(np.float32)


def _geometry(world):
    l1, l2, l3 = FACTORS
    L1, S1 = 1 << l1, 1 << (l2 + l3)
    return L1, S1, L1 // world, S1 // world


def _column_block(x, rank, world):
    L1, S1, rows, cols = _geometry(world)
    return np.ascontiguousarray(x.reshape(L1, S1, 2)[:, rank * cols:(rank + 1) * cols])


def _phase(L, phase, rank, world, src, out_floats):
    out = np.zeros(out_floats, np.float32)
    src = np.ascontiguousarray(src, np.float32).reshape(-1)
    assert L.emu_dist_phase(N_LOG, *FACTORS, phase, rank, world, src.ctypes.data_as(fp), out.ctypes.data_as(fp)) == 0
    return out


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = _emu()
    L1, S1, rows, cols = _geometry(world)
    x = _signal()
    send = _phase(L, 0, rank, world, _column_block(x, rank, world), L1 * cols * 2).reshape(world, rows * cols * 2)
    # all-to-all of equal contiguous chunks: chunk h of `send` goes to rank h, lands as chunk `rank` there
    recv = np.zeros_like(send)
    recv[rank] = send[rank]
    reqs, bufs = [], {}
    for h in range(world):
        if h == rank:
            continue
        bufs[h] = torch.zeros(rows * cols * 2)
        reqs.append(dist.isend(torch.from_numpy(send[h].copy()), dst=h))
        reqs.append(dist.irecv(bufs[h], src=h))
    for r in reqs:
        r.wait()
    for h, b in bufs.items():
        recv[h] = b.numpy()
    nat = _phase(L, 1, rank, world, recv, rows * S1 * 2)
    out_t = _phase(L, 2, rank, world, nat, S1 * rows * 2)
    X = np.fft.fft(x[:, 0].astype(np.float64) + 1j * x[:, 1])
    want = X.reshape(S1, L1)[:, rank * rows:(rank + 1) * rows]
    got = out_t.reshape(S1, rows, 2)
    got = got[..., 0] + 1j * got[..., 1]
    err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    dist.barrier()
    q.put((rank, err))
    dist.destroy_process_group()


def test_two_rank_gloo_distributed_four_step():
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-6 * N_LOG, (rank, err)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_peer_store_phase0_equals_all_to_all(world):
    """fft_dist_phase0_peer's stores (kernel source, emulated) put every row block where the all-to-all of the
    plain phase 0 output would: buffer h, chunk `rank`."""
    L = _emu()
    L1, S1, rows, cols = _geometry(world)
    x = _signal()
    bufs = [np.full((world, rows * cols * 2), np.nan, np.float32) for _ in range(world)]
    ptrs = (fp * world)(*[b.ctypes.data_as(fp) for b in bufs])
    sends = []
    for rank in range(world):
        blk = _column_block(x, rank, world).reshape(-1)
        sends.append(_phase(L, 0, rank, world, blk, L1 * cols * 2).reshape(world, rows * cols * 2))
        assert L.emu_dist_phase0_peer(N_LOG, *FACTORS, rank, world, blk.ctypes.data_as(fp), ptrs) == 0
    for h in range(world):
        for rank in range(world):
            assert np.array_equal(bufs[h][rank], sends[rank][h]), (h, rank)


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("chunk_rows", [0, 8, 24])
def test_dist_schedule_chunked_and_natural_order(world, chunk_rows):
    """fft_dist_transform's schedule for phases 1 + 2 (large_plan.h: build_dist_schedule), kernel source emulated:
    whole-array or L2-chunked (ragged last chunk), transposed-out == the plain phases; natural order with the second
    all-to-all fused into pass C's stores == the contiguous blocks of the spectrum on every rank."""
    L = _emu()
    L1, S1, rows, cols = _geometry(world)
    N = 1 << N_LOG
    x = _signal()
    X = np.fft.fft(x[:, 0].astype(np.float64) + 1j * x[:, 1])
    # exchange-layout buffers of every rank, produced by the (already tested) peer-store phase 0
    bufs = [np.zeros((world, rows * cols * 2), np.float32) for _ in range(world)]
    ptrs = (fp * world)(*[b.ctypes.data_as(fp) for b in bufs])
    for rank in range(world):
        blk = _column_block(x, rank, world).reshape(-1)
        assert L.emu_dist_phase0_peer(N_LOG, *FACTORS, rank, world, blk.ctypes.data_as(fp), ptrs) == 0
    chunk_elems = chunk_rows * S1
    nats = [np.full(2 * N // world, np.nan, np.float32) for _ in range(world)]
    nat_ptrs = (fp * world)(*[b.ctypes.data_as(fp) for b in nats])
    for rank in range(world):
        recv = bufs[rank].reshape(-1)
        plain = _phase(L, 2, rank, world, _phase(L, 1, rank, world, recv, rows * S1 * 2), S1 * rows * 2)
        out_t = np.full(S1 * rows * 2, np.nan, np.float32)
        nl = L.emu_dist_schedule(N_LOG, *FACTORS, rank, world, 0, chunk_elems, 2, recv.ctypes.data_as(fp), out_t.ctypes.data_as(fp), nat_ptrs)
        assert nl >= 2
        if chunk_rows and chunk_rows < rows:
            assert nl == 2 * -(-rows // chunk_rows)
        assert np.array_equal(out_t, plain)
        assert L.emu_dist_schedule(N_LOG, *FACTORS, rank, world, 1, chunk_elems, 3, recv.ctypes.data_as(fp), None, nat_ptrs) >= 2
    for rank in range(world):
        got = nats[rank].reshape(-1, 2)
        got = got[:, 0].astype(np.float64) + 1j * got[:, 1]
        want = X[rank * N // world:(rank + 1) * N // world]
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-6 * N_LOG
