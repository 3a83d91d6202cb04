"""Multi-GPU sharding of independent draws: one process per GPU (torchrun), static
contiguous blocks of draws per rank, no collective on the data path.  The only exchanges
are the final gather of fit parameters (a few KB per draw) and, for the time-mean of
compute_psf_from_sparta (psfrec.py:1104), one sum of [nl, 40, 40] per rank.

This replaces the reference's joblib process pool (psfrec.py:1082-1083).
"""
import numpy as np


def partition(n_items, world, rank):
    """Contiguous block [start, stop) of rank `rank`; blocks differ by at most one item."""
    base, rem = divmod(int(n_items), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def gather_blocks(local, n_items, dst=0):
    """Gather per-rank blocks (first axis = this rank's draws) to rank `dst` in draw order.
    Returns the assembled numpy array on `dst`, None elsewhere.  Works with gloo (CPU
    tensors) and nccl (tensors are staged on the current CUDA device)."""
    import torch
    dist = _dist()
    local = np.ascontiguousarray(local)
    if dist is None or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    use_cuda = dist.get_backend() == 'nccl'
    dev = torch.device('cuda', torch.cuda.current_device()) if use_cuda else torch.device('cpu')
    sizes = [partition(n_items, world, r) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad = np.zeros((width,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    mine = torch.from_numpy(pad).to(dev)
    bufs = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, bufs, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([bufs[r].cpu().numpy()[:b - a] for r, (a, b) in enumerate(sizes)], axis=0)


def allreduce_sum(arr):
    """Sum a small array over ranks (time-mean of the PSF cubes)."""
    import torch
    dist = _dist()
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    if dist is None or dist.get_world_size() == 1:
        return arr
    use_cuda = dist.get_backend() == 'nccl'
    t = torch.from_numpy(arr.copy())
    if use_cuda:
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def compute_psf_sharded(lbda, seeing, GL, L0, h=(100, 10000), npsflin=1, three_lgs_mode=False,
                        want_cube=False, compute_fn=None):
    """compute_psf_batch over the ranks of the current process group.

    Every rank passes the FULL parameter arrays, processes its own block on its own GPU and
    rank 0 receives (fit [ndraw, nl, 16], cube or None, cube_sum [nl, 40, 40]); the other ranks
    receive (None, None, cube_sum).  `compute_fn` (tests only) replaces the CUDA call."""
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    seeing, GL, L0 = (np.atleast_1d(np.asarray(v, dtype=float)) for v in (seeing, GL, L0))
    n = seeing.size
    a, b = partition(n, world, rank)
    h_arr = np.array(h)
    h_loc = h_arr[a:b] if h_arr.ndim == 2 else h_arr
    if compute_fn is None:
        from . import psfrec
        compute_fn = psfrec.compute_psf_batch
    lam = np.atleast_1d(np.asarray(lbda, dtype=float))
    if b > a:
        fit, cube = compute_fn(lam, seeing[a:b], GL[a:b], L0[a:b], npsflin=npsflin, h=h_loc,
                               three_lgs_mode=three_lgs_mode)
    else:
        fit, cube = np.zeros((0, lam.size, 16)), np.zeros((0, lam.size, 40, 40))
    cube_sum = allreduce_sum(cube.sum(axis=0))
    fit_all = gather_blocks(fit, n)
    cube_all = gather_blocks(cube, n) if want_cube else None
    return fit_all, cube_all, cube_sum
