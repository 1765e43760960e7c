"""World-size-2 checks of the multi-GPU plumbing on CPU (gloo): time-shard ranges, the filter-width halo
exchange of the injected DF mode and the all-reduce of the partial Gram matrices.  The kernels are replaced
by the CPU oracle here -- what is under test is the sharding / collective logic of pipeline.py."""
import os
import socket

import numpy as np
import pytest

from tests.util import pkg

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_shard_range_partitions():
    pl = pkg().pipeline
    for n in (20000, 17, 8, 3):
        for world in (1, 2, 3, 8):
            cuts = [pl.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
            cuts8 = [pl.shard_range(n, r, world, 8) for r in range(world)]
            assert cuts8[0][0] == 0 and cuts8[-1][1] == n
            assert all(cuts8[i][1] == cuts8[i + 1][0] for i in range(world - 1))
            assert all(a % 8 == 0 for a, b in cuts8 if b > a)      # every non-empty shard starts on a multiple of 8


def _df_halo(rank, world):
    """Each rank holds its own slabs of one global stream, fetches the 2*nfx-slab halo from its right
    neighbour, filters its time shard with the oracle; rank 0 checks the concatenation against one pass."""
    from oracle import df_oracle as dfo
    pl = pkg().pipeline
    J, K, nf, ln, ns = 6, 7, 2, 1.0, 13
    X = dfo.reference_stream(7, ns, J, K, nf, nf, nf)[:, :ns + 2 * nf]          # (3, ns+2nf, Jp, Kp)
    slabs_all = torch.from_numpy(np.ascontiguousarray(np.moveaxis(X, 0, 1)))    # (ns+2nf, 3, Jp, Kp)
    t0, t1 = pl.shard_range(ns, rank, world)
    own_hi = t1 if rank < world - 1 else ns + 2 * nf                             # last rank keeps the tail
    mine = slabs_all[t0:own_hi].clone()
    win = pl.exchange_halo(mine, 2 * nf)
    assert win.shape[0] == (t1 - t0) + 2 * nf
    assert torch.equal(win, slabs_all[t0:t1 + 2 * nf])
    A = dfo.generate(np.moveaxis(win.numpy(), 1, 0), t1 - t0, J, K, ln, ln, ln, nf, nf, nf, None)
    parts = [None] * world
    dist.all_gather_object(parts, A)
    whole = dfo.generate(X, ns, J, K, ln, ln, ln, nf, nf, nf, None)
    return bool(np.array_equal(np.concatenate(parts, axis=1), whole))


def test_df_time_shards_with_halo_exchange():
    assert all(_run(_df_halo, 2))
    assert all(_run(_df_halo, 3))


def _gram_allreduce(rank, world):
    """Row-sharded Gram: per-rank partial (oracle) + one all-reduce == the full correlation matrix."""
    from oracle import pod_oracle as podo
    pl = pkg().pipeline
    M, N = 3 * 40, 9
    A = podo.synthetic_snapshots(M, N, seed=3, rank_r=5)
    A = A - A.mean(1)[:, None]
    m0, m1 = pl.shard_range(M, rank, world)
    part = torch.from_numpy(A[m0:m1].T @ A[m0:m1] / N)
    pl.allreduce_gram(part)
    full = np.zeros((N, N))
    podo.calculate_correlation_matrix(N, M // 3, 3, "false", [], A, full)
    ok = np.abs(part.numpy() - full).max() < 1e-14 * np.abs(full).max()
    # replicated eigen post-processing gives the same modes on every rank
    lam, V = np.linalg.eigh(part.numpy())
    energy, nm, T = pl.select_modes(lam, lambda j: V[:, j], N, 4)
    ref = podo.pod(A, N, M // 3, 3, "false", [], 1.0e-15, 4, eig=np.linalg.eigh)
    ok = ok and nm == ref["nm"] and np.abs(energy[:nm] - ref["energy"][:nm].real).max() < 1e-12
    sign = np.sign(np.sum(T * ref["temporal_modes"][:, :nm].real, axis=0))
    ok = ok and np.abs(T * sign - ref["temporal_modes"][:, :nm].real).max() < 1e-9
    return bool(ok)


def test_pod_row_shards_allreduce():
    assert all(_run(_gram_allreduce, 2))


def _reshard(rank, world):
    """Time-sharded planes -> row-sharded snapshot matrix with one all-to-all."""
    pl = pkg().pipeline
    N, M = 21, 3 * 17
    full = torch.arange(N * M, dtype=torch.float64).reshape(N, M)
    t0, t1 = pl.shard_range(N, rank, world, 8)
    a_local, (m0, m1) = pl.reshard_time_to_rows(full[t0:t1].clone())
    return bool(torch.equal(a_local, full[:, m0:m1]))


def test_df_to_pod_reshard_all_to_all():
    assert all(_run(_reshard, 2))
    assert all(_run(_reshard, 3))
