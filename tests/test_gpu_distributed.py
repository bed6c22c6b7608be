"""Distributed four-step transform (fft_dist_phase + NCCL all-to-all).  world = 1 runs on any GPU box;
world = 2 needs two devices (gpurun --gpus 2) and is skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _signal(N):
    g = torch.Generator(device="cpu").manual_seed(42)
    return torch.rand(N, 2, generator=g) * 2 - 1  # every rank regenerates the same global signal


def _run_rank(rank, world, port, n, q, exchange="peer"):
    import torch.distributed as dist

    from chowdsp_fft_b200.distributed import DistributedFFT, column_block

    torch.cuda.set_device(rank)
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    N = 1 << n
    x = _signal(N)
    d = DistributedFFT(n, rank, world, exchange=exchange)
    xc = column_block(x.cuda(), d.L1, d.S1, rank, world)
    out_t = torch.empty(d.S1 * d.rows * 2, device="cuda")
    d.forward(xc, out_t)
    nat = d.natural(out_t)
    torch.cuda.synchronize()
    if exchange == "peer":
        # natural order with the second all-to-all fused into the last pass's peer stores == the NCCL redistribution, bit for bit
        nat2 = torch.empty_like(nat)
        d.forward(xc, nat2, natural=True, timed=True)
        torch.cuda.synchronize()
        assert torch.equal(nat2, nat)
        assert len(d.phase_ms()) == 4 and all(t >= 0 for t in d.phase_ms())
    X = np.fft.fft(x[:, 0].double().numpy() + 1j * x[:, 1].double().numpy())
    want_t = X.reshape(d.S1, d.L1)[:, rank * d.rows:(rank + 1) * d.rows]
    got_t = out_t.view(d.S1, d.rows, 2).cpu().numpy()
    got_t = got_t[..., 0] + 1j * got_t[..., 1]
    e1 = np.linalg.norm(got_t - want_t) / np.linalg.norm(want_t)
    got_n = nat.view(-1, 2).cpu().numpy()
    got_n = got_n[:, 0] + 1j * got_n[:, 1]
    want_n = X[rank * N // world:(rank + 1) * N // world]
    e2 = np.linalg.norm(got_n - want_n) / np.linalg.norm(want_n)
    # backward of the same data layout returns N * conj-symmetric partner: check against ifft
    d.forward(xc, out_t, direction=1)
    torch.cuda.synchronize()
    Xi = np.fft.ifft(x[:, 0].double().numpy() + 1j * x[:, 1].double().numpy()) * N
    got_b = out_t.view(d.S1, d.rows, 2).cpu().numpy()
    got_b = got_b[..., 0] + 1j * got_b[..., 1]
    want_b = Xi.reshape(d.S1, d.L1)[:, rank * d.rows:(rank + 1) * d.rows]
    e3 = np.linalg.norm(got_b - want_b) / np.linalg.norm(want_b)
    d.close()
    if world > 1:
        dist.destroy_process_group()
    if q is not None:
        q.put((rank, float(e1), float(e2), float(e3)))
    return e1, e2, e3


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("n", [21, 23])
def test_dist_phases_world1(n, exchange):
    tol = 1e-6 * n
    e1, e2, e3 = _run_rank(0, 1, 0, n, None, exchange)
    assert e1 < tol and e2 < tol and e3 < tol, (e1, e2, e3)


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_dist_multi_gpu(world, exchange):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    import torch.multiprocessing as mp

    n = 22 if world == 2 else 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_rank, args=(r, world, port, n, q, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert max(r[1:]) < 1e-6 * n, r


def _hash_signal(idx):
    """Counter-based synthetic signal keyed by the global FLOAT index (every rank / layout regenerates the same values):
    64-bit LCG step, top bits -> U(-1, 1).  idx: int64 tensor."""
    h = idx * 6364136223846793005 + 1442695040888963407
    h = (h ^ (h >> 29)) * 2862933555777941757 + 3037000493
    return ((h >> 40) & 0xFFFFFF).double() / float(1 << 23) - 1.0


def _sampled_bins(N, ks, device):
    """float64 DFT bins of the hash signal, exact integer phases, chunked"""
    out = []
    step = 1 << 23
    for k in ks:
        re = im = 0.0
        for n0 in range(0, N, step):
            n = torch.arange(n0, n0 + step, device=device, dtype=torch.int64)
            ang = ((n * int(k)) % N).double() * (-2.0 * np.pi / N)
            c, s_ = torch.cos(ang), torch.sin(ang)
            xr, xi = _hash_signal(2 * n), _hash_signal(2 * n + 1)
            re += float((xr * c - xi * s_).sum())
            im += float((xr * s_ + xi * c).sum())
        out.append(complex(re, im))
    return np.array(out)


def _run_rank_full_size(rank, world, port, q):
    """BASELINE configs[4] at its stated size: N = 2^28 over `world` GPUs; sampled bins of this rank's block against a
    float64 DFT, for the transposed-out and the natural-order contract."""
    import torch.distributed as dist

    from chowdsp_fft_b200.distributed import DistributedFFT

    torch.cuda.set_device(rank)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n = 28
    N = 1 << n
    d = DistributedFFT(n, rank, world, exchange="peer")
    n1 = torch.arange(d.L1, device="cuda", dtype=torch.int64)[:, None]
    c = torch.arange(d.cols, device="cuda", dtype=torch.int64)[None, :]
    g = n1 * d.S1 + rank * d.cols + c                               # global complex index of every element of the column block
    xc = torch.stack([_hash_signal(2 * g).float(), _hash_signal(2 * g + 1).float()], dim=-1).contiguous()
    del g
    out_t = torch.empty(d.S1 * d.rows * 2, device="cuda")
    d.forward(xc, out_t)
    nat = torch.empty(2 * (N // world), device="cuda")
    d.forward(xc, nat, natural=True)
    torch.cuda.synchronize()
    rng = np.random.default_rng(100 + rank)
    # transposed-out: out[q][k] = X[(rank*rows + k) + L1*q]
    qs, ks = rng.integers(0, d.S1, 6), rng.integers(0, d.rows, 6)
    bins = [int(rank * d.rows + k + d.L1 * qq) for qq, k in zip(qs, ks)]
    want = _sampled_bins(N, bins, "cuda")
    got = out_t.view(d.S1, d.rows, 2)[torch.tensor(qs, device="cuda"), torch.tensor(ks, device="cuda")].cpu().numpy()
    got = got[:, 0].astype(np.float64) + 1j * got[:, 1]
    e1 = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    # natural order: block[i] = X[rank*N/world + i]
    idx = rng.integers(0, N // world, 6)
    want = _sampled_bins(N, [int(rank * (N // world) + i) for i in idx], "cuda")
    got = nat.view(-1, 2)[torch.tensor(idx, device="cuda")].cpu().numpy()
    got = got[:, 0].astype(np.float64) + 1j * got[:, 1]
    e2 = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    d.close()
    dist.destroy_process_group()
    q.put((rank, e1, e2, 0.0))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_config5_full_size_distributed(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_rank_full_size, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert max(r[1:]) < 1e-6 * 28, r
