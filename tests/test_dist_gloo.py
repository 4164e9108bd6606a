"""CPU, world_size 2 over gloo: the sharding + all-gather plumbing of the multi-GPU path.

The exchange only moves packed per-rank states ((p+1)^2 + 2(p+1) doubles: R, column norms,
row count, column order).  Here each
rank builds its packed state with the CPU oracle (numpy QR of its shard), the product's
all-gather helper moves them, and the merged factor must reproduce the global index sets
and OLS solution.  The CUDA merge kernel itself is covered by the -m gpu tests
(test_gpu_multi.py: simulated shards on one GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _pack(W, tau):
    p = W.shape[1]
    A = np.c_[W, tau]
    R = np.linalg.qr(A, mode="r")
    Rfull = np.zeros((p + 1, p + 1))
    Rfull[:R.shape[0]] = R
    return np.concatenate([Rfull.reshape(-1), (A * A).sum(0), [float(W.shape[0])], np.arange(p, dtype=float)])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "figaroh-plus_b200")):
        sys.path.insert(0, p)
    from figaroh_b200.dist import allgather_packed, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)               # same global matrix on every rank
        M, p = 901, 12
        W = rng.normal(size=(M, p))
        W[:, 7] = W[:, 2] - 2 * W[:, 5]              # a dependent column
        W[:, 3] = 0.0                                # a structurally zero column
        tau = W @ rng.normal(size=p) + 0.01 * rng.normal(size=M)
        lo, hi = shard_range(M, rank, world)
        packed = torch.from_numpy(_pack(W[lo:hi], tau[lo:hi]))
        allp = allgather_packed(packed).numpy().reshape(world, -1)
        p1 = p + 1
        Rs = [allp[r, :p1 * p1].reshape(p1, p1) for r in range(world)]
        colsq = sum(allp[r, p1 * p1:p1 * p1 + p1] for r in range(world))
        rows = sum(allp[r, p1 * p1 + p1] for r in range(world))
        ok0 = all(np.array_equal(allp[r, p1 * p1 + p1 + 1:], np.arange(p)) for r in range(world))
        Rm = np.linalg.qr(np.vstack(Rs), mode="r")
        A = np.c_[W, tau]
        ok = (rows == M) and ok0
        ok &= np.allclose(colsq, (A * A).sum(0), rtol=1e-12)
        ok &= np.allclose(Rm.T @ Rm, A.T @ A, rtol=1e-10, atol=1e-9)
        keep = [i for i in range(p) if colsq[i] >= 1e-6]
        ok &= keep == [i for i in range(p) if i != 3]
        Re = np.linalg.qr(Rm[:, keep + [p]], mode="r")
        base = [i for i in range(len(keep)) if abs(Re[i, i]) > 1e-8]
        ok &= base == [i for i in range(len(keep)) if keep[i] != 7]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_allgather_of_packed_states_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
