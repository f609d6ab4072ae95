"""Multi-GPU post-correction: one process per GPU (torch.distributed), draws
sharded in contiguous static blocks -- the reference's `#pragma omp for
schedule(static)` over draws (/root/reference/src/approx_mcmc.cpp:364-369) --
and ONE all_gather of the per-draw results at the end.  No collective on the
data path: filters are independent.  Because draw d always uses random stream
`filter = d`, the gathered result is bit-identical for any number of ranks."""
import numpy as np


def shard_range(n_draws, rank, world):
    """contiguous static block of rank `rank` out of `world`"""
    return (n_draws * rank) // world, (n_draws * (rank + 1)) // world


def gather_shards(local, n_draws, rank, world, device=None):
    """all_gather of a 1-D or 2-D float64 array whose LAST axis is the draw axis;
    shards may differ in length by one (padded for the collective)."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, dtype=np.float64)
    lead = local.shape[:-1]
    width = int(np.prod(lead)) if lead else 1
    maxlen = max(shard_range(n_draws, r, world)[1] - shard_range(n_draws, r, world)[0] for r in range(world))
    buf = torch.zeros(width * maxlen, dtype=torch.float64, device=device)
    flat = torch.from_numpy(local.reshape(width, -1)).to(buf.device)
    buf.view(width, maxlen)[:, :flat.shape[1]] = flat
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    parts = []
    for r in range(world):
        b, e = shard_range(n_draws, r, world)
        parts.append(outs[r].view(width, maxlen)[:, :e - b].cpu().numpy())
    return np.concatenate(parts, axis=1).reshape(lead + (n_draws,))


def post_correct_distributed(model, mcmc_output, particles, is_type="is2", seed=1, compute=None, device=None):
    """post_correct over all ranks of the default process group.  `compute` is the
    per-shard function (default: bssm_b200.post_correct on this rank's GPU)."""
    import torch.distributed as dist
    from . import api
    rank, world = dist.get_rank(), dist.get_world_size()
    it = {"is1": 1, "is2": 2, "is3": 3}[is_type.lower()]
    total = api.n_draws(mcmc_output, it)
    if int(mcmc_output.get("output_type", 3)) == 2 and world > 1:
        # the summary merge (weighted running mean / variance over draws, src/approx_mcmc.cpp:397-424) is
        # finished inside the C call (Vt += Valpha / sum_w), so per-rank summaries cannot be combined here;
        # bssm_b200_postcorrection_* with n_threads > 1 does the multi-GPU merge inside the library
        raise ValueError("post_correct_distributed: output_type 'summary' is not shardable across processes; "
                         "call post_correct(..., threads=G) instead (one process, G GPUs)")
    b, e = shard_range(total, rank, world)
    fn = compute or api.post_correct
    local = fn(model, mcmc_output, particles, is_type=is_type, seed=seed, draw_range=(b, e))
    res = dict(local)
    res["weights"] = gather_shards(local["weights"], total, rank, world, device)
    res["posterior"] = gather_shards(local["posterior"], total, rank, world, device)
    if int(mcmc_output.get("output_type", 3)) == 1:
        res["alpha"] = gather_shards(local["alpha"], total, rank, world, device)
    return res
