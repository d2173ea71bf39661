"""CPU suite for the N > 1 path (world_size 2, gloo backend, 127.0.0.1): chain / row partitions, the row-sharding exchange
step (all-reduce of log-likelihood partials, priors added once) against the oracle's full-data value, the all-reduce hook
with the C-ABI callback signature, and the ESS aggregation used by bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, fn, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        out[rank] = fn(rank)
    finally:
        dist.destroy_process_group()


def _run(fn):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(_free_port(), fn, out), nprocs=WORLD, join=True)
    return [out[r] for r in range(WORLD)]


def test_partitions_cover_and_are_disjoint():
    from cmdstan_b200 import sharding as S
    for world in (1, 2, 3, 8):
        for n in (1, 7, 64, 1000, 10_000_001):
            rows = [S.row_shard(r, world, n) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == n and all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in rows]
            assert max(sizes) - min(sizes) <= 1
        ch = [S.chain_shard(r, world, 8192, init_chain_id=5) for r in range(world)]
        ids = [i for first, n, fid in ch for i in range(fid, fid + n)]
        assert ids == list(range(5, 5 + 8192))   # stream ids: contiguous, no overlap, nothing missing


def _row_sharded_glm(rank):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cmdstan_b200 import sharding as S
    from oracle import oracle as O
    res = {}
    for name, data in (("logistic", cases.logistic_data(301, 6)), ("poisson_glm", cases.poisson_data(257, 5))):
        q = np.random.default_rng(3).uniform(-1, 1, size=(4, data["X"].shape[1] + 1))
        local = S.shard_rows(data, rank, WORLD)
        prior = O.Model(name, {"X": data["X"][:0], "y": data["y"][:0]})
        lp0, g0 = prior.lp_grad(q)                      # zero rows: the prior terms only
        lpl, gl = O.Model(name, local).lp_grad(q)
        lp, g = S.combine_row_partials(lpl - lp0, gl - g0, lp0, g0)   # partial log-likelihood of this rank's rows
        res[name] = (lp.numpy(), g.numpy())
    return res


def test_row_sharding_exchange_equals_full_data_oracle():
    from oracle import oracle as O
    out = _run(_row_sharded_glm)
    for name, data in (("logistic", cases.logistic_data(301, 6)), ("poisson_glm", cases.poisson_data(257, 5))):
        q = np.random.default_rng(3).uniform(-1, 1, size=(4, data["X"].shape[1] + 1))
        lp_full, g_full = O.Model(name, data).lp_grad(q)
        for r in range(WORLD):
            lp, g = out[r][name]
            assert np.allclose(lp, lp_full, rtol=1e-12) and np.allclose(g, g_full, rtol=1e-11, atol=1e-11)
        # bit-identical on every rank: identical control flow downstream (SURVEY.md section 8e "determinism")
        assert np.array_equal(out[0][name][0], out[1][name][0]) and np.array_equal(out[0][name][1], out[1][name][1])


def _hook_roundtrip(rank):
    import ctypes as C
    from cmdstan_b200 import ALLREDUCE_FN, sharding as S
    hook = ALLREDUCE_FN(S.torch_allreduce_hook(None, "cpu"))   # the C-ABI callback type of b200nuts_set_allreduce
    buf = np.arange(10, dtype=np.float64) * (rank + 1)
    rc = hook(buf.ctypes.data_as(C.c_void_p), 10, None, None)
    return rc, buf.tolist()


def test_allreduce_hook_c_signature():
    out = _run(_hook_roundtrip)
    want = (np.arange(10) * 3.0).tolist()
    assert out[0] == (0, want) and out[1] == (0, want)


def _ess_sum(rank):
    from cmdstan_b200 import sharding as S
    t = torch.tensor([100.0 + rank, 50.0, 7.0 * (rank + 1)], dtype=torch.float64)
    return S.sum_over_ranks(t).tolist()


def test_ess_aggregation_over_chain_shards():
    out = _run(_ess_sum)
    assert out[0] == out[1] == [201.0, 100.0, 21.0]
