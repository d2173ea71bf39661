"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing) -- SURVEY.md section 8e.

Two ways the path shards, both without changing any chain's output:

  * chains   (c1-c4): rank r runs chains [r*C/G, (r+1)*C/G) with stream ids init_chain_id + first_chain; chains are
             independent (the reference's own parallel-match invariant,
             stan/src/test/unit/services/sample/hmc_nuts_diag_e_adapt_parallel_match_test.cpp:73-153), so there is NO
             data-path collective; only the ESS vectors are summed at the end.
  * rows     (c5): every rank holds all chains with identical RNG streams and rows [r*N/G, (r+1)*N/G) of the design
             matrix; the log-likelihood is a sum over rows, so each leapfrog needs exactly one exchange step:
             all-reduce(sum) of the (lp, gradient) partials [(1 + D) doubles per chain]; priors are added once, after it.
             Product path: ncclAllReduce issued natively by libb200nuts on the engine's stream (Engine.nccl_init).
             `torch_allreduce_hook` is the same exchange through torch.distributed (any backend) -- used for the
             world-size-2 gloo tests on CPU and available as a fallback transport.
"""
import ctypes as C

import numpy as np


def chain_shard(rank, world, num_chains, init_chain_id=1):
    """-> (first_chain, n_local, local_init_chain_id): contiguous, balanced, covering, disjoint."""
    first = num_chains * rank // world
    last = num_chains * (rank + 1) // world
    return first, last - first, init_chain_id + first


def row_shard(rank, world, n_rows):
    """-> (row_begin, row_end) of this rank: contiguous, sizes differ by at most one."""
    return n_rows * rank // world, n_rows * (rank + 1) // world


def shard_rows(data, rank, world):
    """Row slice of a GLM data set {"X": [N, K], "y": [N]}."""
    b, e = row_shard(rank, world, len(data["y"]))
    return {"X": np.ascontiguousarray(data["X"][b:e]), "y": np.ascontiguousarray(data["y"][b:e])}


class _DevicePtr:
    """zero-copy view of `count` doubles at a CUDA device address for torch.as_tensor"""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def torch_allreduce_hook(group=None, device=None):
    """All-reduce callback with the signature of b200nuts_allreduce_fn (include/b200nuts.h) on top of torch.distributed.

    device None / "cpu": `buf` is a host pointer (gloo; tests).  device = torch.device("cuda", i): `buf` is a device
    pointer and the collective is ordered on the engine's stream through an ExternalStream."""
    import torch
    import torch.distributed as dist

    def hook(buf, count, stream, user):
        try:
            if device is None or str(device) == "cpu":
                a = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double)), shape=(count,))
                t = torch.from_numpy(a)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:
                t = torch.as_tensor(_DevicePtr(buf, count), device=device)
                with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=device)):
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception:  # never unwind through the C caller
            return 70

    return hook


def broadcast_nccl_id(src=0, group=None):
    """Rank `src` creates the 128-byte NCCL unique id, every rank receives it (torch.distributed object broadcast)."""
    import torch.distributed as dist
    from . import nccl_unique_id
    box = [nccl_unique_id() if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def row_sharded_engine(model, data, rank, world, transport="nccl", group=None, device=0, **engine_kw):
    """Engine over this rank's rows of `data`; all ranks must pass the same chains / seed / ids (identical control flow)."""
    import torch
    from . import Engine
    local = shard_rows(data, rank, world)
    if world == 1:
        return Engine(model, local, device=device, **engine_kw)
    if transport == "nccl":
        e = Engine(model, local, device=device, row_shards=world, row_shard_rank=rank, **engine_kw)
        e.nccl_init(broadcast_nccl_id(group=group), world, rank)
        return e
    return Engine(model, local, device=device, row_shards=world, row_shard_rank=rank,
                  allreduce=torch_allreduce_hook(group, torch.device("cuda", device)), **engine_kw)


def chain_sharded_engine(model, data, rank, world, num_chains, init_chain_id=1, device=0, **engine_kw):
    """Engine over this rank's chains (data replicated)."""
    from . import Engine
    _, n_local, first_id = chain_shard(rank, world, num_chains, init_chain_id)
    return Engine(model, data, num_chains=n_local, init_chain_id=first_id, device=device, **engine_kw)


def sum_over_ranks(t, group=None):
    """ESS of disjoint chain sets adds (independent chains): element-wise sum of a tensor over the ranks, in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def combine_row_partials(lp_partial, grad_partial, prior_lp, prior_grad, group=None):
    """The row-sharding identity on host tensors: given each rank's log-likelihood partials (lp [n], grad [n, D], WITHOUT
    priors), all-reduce them and add the row-independent prior terms once.  Mirrors what glm_reduce -> all-reduce ->
    glm_finish do on the device."""
    import torch
    buf = torch.cat([torch.as_tensor(lp_partial, dtype=torch.float64).reshape(-1, 1), torch.as_tensor(grad_partial, dtype=torch.float64)], dim=1).contiguous()
    sum_over_ranks(buf, group)
    return buf[:, 0] + torch.as_tensor(prior_lp), buf[:, 1:] + torch.as_tensor(prior_grad)
