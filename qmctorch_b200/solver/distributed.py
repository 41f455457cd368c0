"""Multi-GPU plumbing: walkers are sharded across ranks (one process per GPU); the only
collectives are SUM all-reduces of the energy statistics and of the parameter gradients
(SURVEY.md section 8e).  Works with any torch.distributed backend (nccl on GPUs, gloo in
the CPU tests)."""
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world():
    if is_distributed():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_walkers(nwalkers, rank=None, world_size=None):
    """Contiguous split of ``nwalkers`` over the ranks -> (first, count) of this rank."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(nwalkers), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def allreduce_sum_(t):
    """In-place SUM all-reduce (no-op on one rank)."""
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_max_(t):
    """In-place MAX all-reduce (no-op on one rank)."""
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def rank_seed(seed):
    """Folds the rank into a 63-bit seed so that shards draw independent streams even when every rank
    called the same torch.manual_seed (usual DDP practice); rank 0 / single process: unchanged."""
    rank, _ = world()
    return (int(seed) + rank * 0x9E3779B97F4A7C15) & 0x7FFFFFFFFFFFFFFF


def global_stats(sum_e, sum_e2, n, nbad=0.0, device=None):
    """(mean, unbiased variance, standard error, n, nbad) from per-rank partial sums; one
    all-reduce of four doubles."""
    if torch.is_tensor(sum_e) and sum_e.numel() == 4:
        buf = sum_e.clone()
    else:
        buf = torch.tensor([float(sum_e), float(sum_e2), float(n), float(nbad)], dtype=torch.float64,
                           device=device)
    allreduce_sum_(buf)
    s, s2, cnt, bad = (float(v) for v in buf.tolist())
    mean = s / cnt
    var = (s2 - cnt * mean * mean) / (cnt - 1) if cnt > 1 else float("nan")
    err = (var / cnt) ** 0.5 if var == var and var >= 0 else float("nan")
    return mean, var, err, cnt, bad


def allreduce_gradients(params):
    """SUM the ``.grad`` of every parameter over the ranks with ONE flat all-reduce."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not is_distributed():
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off: off + n].view_as(g))
        off += n
