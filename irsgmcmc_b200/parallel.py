"""
Multi-GPU: chains shard by rank with no per-iteration communication.  The only collective is the
merge of the per-rank Welford posterior moments -- two sum all-reduces (NCCL over NVLink on GPUs; gloo in CPU tests).

The reference has no distributed code at all; the quantity computed here is its calc_posterior_statistics
(utils/util.py:114-120: mean and unbiased std over every kept sample of every chain).
"""
import torch
import torch.distributed as dist


def chain_shard(no_chains_total, rank=None, world_size=None):
    """contiguous block of global chain ids owned by `rank`: [offset, offset + count)"""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    base, extra = divmod(no_chains_total, world_size)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def merge_moments(n_local, moments, group=None):
    """
    Chan et al. parallel merge of Welford triples.  `moments` is a list of (mean, M2) tensor pairs that share the
    sample count n_local.  Returns (n_total, (mean, M2), ...) identical on every rank.
        mean = sum_r n_r mean_r / N ;   M2 = sum_r [ M2_r + n_r (mean_r - mean)^2 ]
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return (n_local,) + tuple((m.clone(), m2.clone()) for m, m2 in moments)

    device = moments[0][0].device
    n = torch.tensor([float(n_local)], device=device, dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    n_total = float(n.item())

    # one flat buffer per phase so that each phase is a single collective (bucketed for launch latency, not link count)
    sizes = [m.numel() for m, _ in moments]
    flat = torch.cat([(m.double() * n_local).reshape(-1) for m, _ in moments])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    means = [(part / n_total) for part in flat.split(sizes)]

    flat2 = torch.cat([(m2.double().reshape(-1) + n_local * (m.double().reshape(-1) - mg) ** 2)
                       for (m, m2), mg in zip(moments, means)])
    dist.all_reduce(flat2, op=dist.ReduceOp.SUM, group=group)
    out = []
    for (m, m2), mg, part in zip(moments, means, flat2.split(sizes)):
        out.append((mg.to(m.dtype).view_as(m), part.to(m2.dtype).view_as(m2)))
    return (int(round(n_total)),) + tuple(out)
