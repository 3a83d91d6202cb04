"""Multi-GPU sharding of the independent (draw x wavelength) work items: one process per GPU
(torchrun), static contiguous blocks, no collective on the data path.  The only exchanges are
the final gather of the fit parameters (a few KB per draw) and, for the time-mean of
compute_psf_from_sparta (psfrec.py:1104), one sum of [nl, 40, 40] per rank.

This replaces the reference's joblib process pool (psfrec.py:1082-1083, gather at :1086-1101).

Partitioning: blocks of draws per rank; when there are fewer draws than ranks
(BASELINE configs 1, 3, 5: one draw) the wavelength axis is split as well and stage A is
replicated (it costs 1/nlam of the work).  Field directions of a draw are never split, so the
direction mean stays on the device.
"""
import numpy as np

from ._lib import FIT_NPAR, PSF_DIM


def partition(n_items, world, rank):
    """Contiguous block [start, stop) of rank `rank`; blocks differ by at most one item."""
    base, rem = divmod(int(n_items), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def grid_of(ndraw, nlam, world):
    """(draw splits, wavelength splits): draws first; wavelengths only when ranks are left over."""
    nsd = max(1, min(int(ndraw), int(world)))
    nsl = max(1, min(int(nlam), int(world) // nsd))
    return nsd, nsl


def block_of(ndraw, nlam, world, rank):
    """((d0, d1), (l0, l1)) of `rank`; ranks beyond the grid get empty blocks."""
    nsd, nsl = grid_of(ndraw, nlam, world)
    if rank >= nsd * nsl:
        return (0, 0), (0, 0)
    rd, rl = divmod(rank, nsl)
    return partition(ndraw, nsd, rd), partition(nlam, nsl, rl)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _comm_device(dist):
    import torch
    if dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def gather_grid(local, ndraw, nlam, dst=0, host_out=None):
    """Gather per-rank blocks `local` [nd_loc, nl_loc, ...] (numpy array or torch tensor, host or
    device) into the full [ndraw, nlam, ...] array on rank `dst` (numpy); None elsewhere.  With
    nccl the blocks travel device to device (NVLink) and rank `dst` does one device -> host copy -
    into `host_out` (a pinned torch tensor of the full shape) when given."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local.cpu().numpy() if hasattr(local, 'cpu') else np.ascontiguousarray(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _comm_device(dist)
    blocks = [block_of(ndraw, nlam, world, r) for r in range(world)]
    tail = tuple(local.shape[2:])
    cap = max((d1 - d0) * (l1 - l0) for (d0, d1), (l0, l1) in blocks)
    width = int(np.prod(tail, dtype=np.int64)) if tail else 1
    mine = torch.zeros(cap * width, dtype=torch.float64, device=dev)
    t = local if hasattr(local, 'data_ptr') else torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    mine[:t.numel()] = t.reshape(-1).to(dev, non_blocking=True)
    bufs = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, bufs, dst=dst)
    if rank != dst:
        return None
    full = torch.empty((ndraw, nlam) + tail, dtype=torch.float64, device=dev)
    for r, ((d0, d1), (l0, l1)) in enumerate(blocks):
        n = (d1 - d0) * (l1 - l0)
        if n:
            full[d0:d1, l0:l1] = bufs[r][:n * width].reshape((d1 - d0, l1 - l0) + tail)
    if host_out is not None:
        host_out.copy_(full, non_blocking=True)
        if full.is_cuda:
            torch.cuda.current_stream().synchronize()
        return host_out.numpy()
    return full.cpu().numpy()


def allreduce_sum(arr):
    """Sum a small array over ranks (time-mean of the PSF cubes)."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return arr.cpu().numpy() if hasattr(arr, 'cpu') else np.ascontiguousarray(arr, dtype=np.float64)
    t = arr if hasattr(arr, 'data_ptr') else torch.from_numpy(np.array(arr, dtype=np.float64))
    t = t.to(_comm_device(dist)).clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def compute_psf_sharded(lbda, seeing, GL, L0, h=(100, 10000), npsflin=1, three_lgs_mode=False,
                        want_cube=False, compute_fn=None, out_cube=None, want_sum=True, fit_host=None,
                        **kwargs):
    """compute_psf_batch over the ranks of the current process group.

    Every rank passes the FULL parameter arrays, processes its own (draw, wavelength) block on its
    own GPU and rank 0 receives (fit [ndraw, nl, 16], cube [ndraw, nl, 40, 40] or None, cube_sum
    [nl, 40, 40]); the other ranks receive (None, None, cube_sum).  With the nccl backend the fit
    records stay on the device until rank 0 has gathered them.  ``out_cube`` (optional, pinned host
    tensor or numpy array [nd_loc, nl_loc, 40, 40]) receives this rank's own cube block - the sharded
    output a caller keeps local when only the fits are gathered.  ``compute_fn`` (tests) replaces the
    CUDA call; extra keyword arguments go to it.  ``want_sum=False`` skips the cube sum (then None); ``fit_host``
    (rank 0: pinned tensor [ndraw, nl, 16]) receives the gathered fit records without a pageable copy."""
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    seeing, GL, L0 = (np.atleast_1d(np.asarray(v, dtype=float)) for v in (seeing, GL, L0))
    lam = np.atleast_1d(np.asarray(lbda, dtype=float))
    n, nl = seeing.size, lam.size
    (a, b), (l0, l1) = block_of(n, nl, world, rank)
    h_arr = np.array(h)
    h_loc = h_arr[a:b] if h_arr.ndim == 2 else h_arr
    on_gpu = compute_fn is None
    if on_gpu:
        from . import psfrec
        compute_fn = psfrec.compute_psf_batch
    if b > a and l1 > l0:
        if on_gpu and dist is not None and dist.get_backend() == 'nccl':
            import torch
            dev = _comm_device(dist)
            kwargs.setdefault('out_fit', torch.empty((b - a, l1 - l0, FIT_NPAR), dtype=torch.float64, device=dev))
            if out_cube is None and (want_cube or want_sum):
                out_cube = torch.empty((b - a, l1 - l0, PSF_DIM, PSF_DIM), dtype=torch.float64, device=dev)
        if on_gpu and out_cube is None and not (want_cube or want_sum):
            kwargs['want_cube'] = False          # fits only: the cubes never leave the device workspace
        if out_cube is not None:
            kwargs['out_cube'] = out_cube
        fit, cube = compute_fn(lam[l0:l1], seeing[a:b], GL[a:b], L0[a:b], npsflin=npsflin, h=h_loc,
                               three_lgs_mode=three_lgs_mode, **kwargs)
    else:
        fit, cube = np.zeros((0, 0, FIT_NPAR)), np.zeros((0, 0, PSF_DIM, PSF_DIM))
    # this rank's share of sum over draws of the cubes, placed at its wavelengths
    cube_sum = None
    if want_sum:
        part = np.zeros((nl, PSF_DIM, PSF_DIM))
        if b > a and l1 > l0:
            local_sum = cube.sum(dim=0) if hasattr(cube, 'data_ptr') else np.asarray(cube).sum(axis=0)
            part[l0:l1] = local_sum.cpu().numpy() if hasattr(local_sum, 'cpu') else local_sum
        cube_sum = allreduce_sum(part)
    fit_all = gather_grid(fit, n, nl, host_out=fit_host if rank == 0 else None)
    cube_all = gather_grid(cube, n, nl) if want_cube else None
    return fit_all, cube_all, cube_sum
