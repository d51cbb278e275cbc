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


def any_rank(flag, device=None, group=None):
    """logical OR of a host flag over all ranks (one tiny MAX all-reduce); the flag itself without torch.distributed"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], device=device if dist.get_backend(group) == 'nccl' else 'cpu', dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item())


def merge_moments(n_local, moments, group=None):
    """
    Chan et al. parallel merge of Welford triples.  `moments` is a list of (mean, M2) tensor pairs that share the
    sample count n_local.  Returns (n_total, (mean, M2), ...) identical on every rank.
        mean = sum_r n_r mean_r / N ;   M2 = sum_r [ M2_r + n_r (mean_r - mean)^2 ]
    Two sum all-reduces of ONE packed fp32 buffer each (the payload in its own precision: 32 MiB per phase at 128^3);
    the sample count rides in the first buffer, so nothing returns to the host between the phases -- the only
    synchronisation is the final read of N.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return (n_local,) + tuple((m.clone(), m2.clone()) for m, m2 in moments)

    ref = moments[0][0]
    sizes = [m.numel() for m, _ in moments]
    total = sum(sizes)
    # element 0 carries n (exact in fp32 up to 2^24 samples per rank; the sum is formed in the all-reduce), padded to 4 floats
    # so that the payload views stay 16-byte aligned
    buf = torch.empty(4 + total, device=ref.device, dtype=torch.float32)
    buf[:4] = float(n_local)
    views, o = [], 4
    for (m, _), k in zip(moments, sizes):
        views.append(buf[o:o + k].view_as(m))
        torch.mul(m, float(n_local), out=views[-1])
        o += k
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    n_total = buf[0]                                    # device scalar
    means = [v / n_total for v in views]

    buf2 = torch.empty(total, device=ref.device, dtype=torch.float32)
    out, o = [], 0
    for (m, m2), mg, k in zip(moments, means, sizes):
        d = m - mg
        part = buf2[o:o + k].view_as(m2)
        torch.addcmul(m2, d, d, value=float(n_local), out=part)
        out.append((mg, part))
        o += k
    dist.all_reduce(buf2, op=dist.ReduceOp.SUM, group=group)
    return (int(round(float(n_total.item()))),) + tuple(out)
