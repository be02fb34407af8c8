"""world_size-2 CPU (gloo) test of the multi-GPU plumbing: clip sharding + the single fp64 all-reduce
of the packed sufficient statistics give the same mean / covariance / FAD as one process
(SURVEY.md §8e).  The per-rank accumulate runs in NumPy here (the CUDA kernel needs a GPU); what is
under test is dist.shard_bounds + dist.allreduce_acc + the finalize algebra of csrc/stats.cu."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from frechet_audio_distance_exported_b200.dist import allreduce_acc, shard_bounds
from oracle import stats, synth


def test_shard_bounds_tile_exactly():
    for n in (0, 1, 7, 100, 12500):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _pack(x: np.ndarray) -> torch.Tensor:
    x = x.astype(np.float64)
    d = x.shape[1]
    acc = np.concatenate([[x.shape[0]], x.sum(0), (x.T @ x).reshape(-1)]) if x.shape[0] else np.zeros(1 + d + d * d)
    return torch.from_numpy(acc)


def _finalize(acc: np.ndarray, d: int):
    n, s1, S = acc[0], acc[1:1 + d], acc[1 + d:].reshape(d, d)
    m = s1 / n
    return m, (S - n * np.outer(m, m)) / (n - 1)


def _worker(rank, world, port, n, d, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    accs = []
    for s in (0, 1):
        x = synth.embedding_set(s, n, d)
        lo, hi = shard_bounds(n, rank, world)
        accs.append(_pack(x[lo:hi]))
    both = torch.cat(accs)
    allreduce_acc(both)
    if rank == 0:
        np.save(out, both.numpy())
    dist.destroy_process_group()


def test_two_rank_stats_equal_single_process(tmp_path):
    n, d = 301, 24
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(2, port, n, d, out), nprocs=2, join=True)
    both = np.load(out)
    half = both.size // 2
    a, b = synth.embedding_set(0, n, d), synth.embedding_set(1, n, d)
    mu1, s1 = _finalize(both[:half], d)
    mu2, s2 = _finalize(both[half:], d)
    r1, rs1 = stats.embd_statistics(a)
    r2, rs2 = stats.embd_statistics(b)
    np.testing.assert_allclose(s1, rs1, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(s2, rs2, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(mu1, a.astype(np.float64).mean(0), rtol=1e-12)
    f2 = stats.frechet_distance(mu1, s1, mu2, s2)
    f1 = stats.frechet_distance(a.astype(np.float64).mean(0), rs1, b.astype(np.float64).mean(0), rs2)  # same fp64 means
    assert abs(f1 - f2) / abs(f1) < 1e-10


def test_allreduce_is_noop_without_process_group():
    t = torch.arange(5, dtype=torch.float64)
    assert torch.equal(allreduce_acc(t.clone()), t)
    with pytest.raises(TypeError):
        allreduce_acc(torch.zeros(3, dtype=torch.float32))
