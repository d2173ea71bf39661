"""Run under torchrun (one rank per GPU): checks of the two sharding modes on real GPUs.  Prints MULTI_GPU_OK on rank 0.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/multi_gpu_worker.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import cmdstan_b200 as B  # noqa: E402
from cmdstan_b200 import sharding as S  # noqa: E402


def run(e, n):
    e.init_chains()
    e.run(n)
    err, it, ng = e.chain_status()
    assert (err == 0).all() and (it == n).all()
    d = e.fetch_draws(constrained=False)
    eps, im = e.adaptation()
    e.close()
    return d, eps


def gather(x, rank, world):
    out = [None] * world
    dist.all_gather_object(out, x)
    return out


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    kw = dict(random_seed=4242, num_warmup=40, num_samples=20, save_warmup=True)
    # ---- rows: poisson GLM, every rank all chains, one all-reduce per leapfrog
    data = cases.poisson_data(4003, 6)   # not a multiple of the world size: shards differ by one row
    res = {}
    for transport in ("nccl", "torch"):
        e = S.row_sharded_engine("poisson_glm", data, rank, world, transport=transport, device=lr, num_chains=5, init_chain_id=1, **kw)
        q = np.random.default_rng(1).uniform(-1, 1, size=(7, e.D))
        lp, g = e.log_prob_grad(q)
        if transport == "nccl":   # the exchange step's own cost is measurable (bench.py reports it per leapfrog)
            us, nbytes = e.measure_allreduce(50)
            # payload: (1 + D padded to 8) doubles per workspace chain slot (>= the 5 chains: log_prob_grad above sized it for 7 points)
            assert nbytes % ((1 + 8) * 8) == 0 and nbytes >= 5 * (1 + 8) * 8 and 0.5 < us < 5000.0, (us, nbytes)
        d, eps = run(e, 60)
        res[transport] = (lp, g, d, eps)
        allr = gather((lp, g, d, eps), rank, world)
        for o in allr[1:]:   # bit-identical on every rank => identical control flow
            assert all(np.array_equal(a, b) for a, b in zip(allr[0], o)), "ranks disagree"
    assert all(np.array_equal(a, b) for a, b in zip(res["nccl"], res["torch"])), "nccl and torch transports disagree"
    if rank == 0:
        e = B.Engine("poisson_glm", data, num_chains=5, init_chain_id=1, device=lr, **kw)
        q = np.random.default_rng(1).uniform(-1, 1, size=(7, e.D))
        lp1, g1 = e.log_prob_grad(q)
        d1, eps1 = run(e, 60)
        lp, g, d, eps = res["nccl"]
        assert np.allclose(lp, lp1, rtol=1e-12) and np.allclose(g, g1, rtol=1e-10, atol=1e-9)
        assert np.array_equal(d[:, :6, 3:6], d1[:, :6, 3:6]) and np.allclose(d[:, :6], d1[:, :6], rtol=1e-8, atol=1e-10)
    # ---- chains: 24 chains over the ranks == one engine with 24 chains, bit for bit (eight schools: fused engine; IRT: lock-step
    # engine + latency kernel, the mode of north_star's config c4)
    ok = True
    for model, mdata in (("eight_schools_nc", cases.EIGHT_SCHOOLS), ("irt_2pl", cases.irt_data(40, 9))):
        e = S.chain_sharded_engine(model, mdata, rank, world, num_chains=24, init_chain_id=3, device=lr, **kw)
        d, eps = run(e, 60)
        parts = gather((d, eps), rank, world)
        if rank == 0:
            e = B.Engine(model, mdata, num_chains=24, init_chain_id=3, device=lr, **kw)
            d1, eps1 = run(e, 60)
            ok = ok and np.array_equal(np.concatenate([p[0] for p in parts]), d1) and np.array_equal(np.concatenate([p[1] for p in parts]), eps1)
    if rank == 0:
        assert ok
        print("MULTI_GPU_OK world=%d" % world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
