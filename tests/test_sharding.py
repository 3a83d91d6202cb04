"""CPU tests of the multi-GPU host logic (world_size 2, gloo): block partition, gather of
the fit parameters in draw order, and the cube sum used by the time mean."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from muse_psfr_b200 import sharding


def fake_compute(lam, seeing, GL, L0, npsflin=1, h=(100, 10000), three_lgs_mode=False):
    """Deterministic stand-in for the CUDA call: values depend only on the (draw, wavelength) inputs."""
    nd, nl = seeing.size, lam.size
    fit = np.zeros((nd, nl, 16))
    fit[:, :, 5] = seeing[:, None] * 4 + lam[None, :] * 1e-3
    fit[:, :, 4] = GL[:, None] + L0[:, None] * 0.01
    h = np.array(h, dtype=float)
    fit[:, :, 0] = (h[:, 0] if h.ndim == 2 else h[0])[..., None] if h.ndim == 2 else h[0]
    cube = np.ones((nd, nl, 40, 40)) * seeing[:, None, None, None] * lam[None, :, None, None]
    return fit, cube


def test_partition_covers_everything():
    for n in (0, 1, 7, 30, 4096):
        for world in (1, 2, 3, 8):
            blocks = [sharding.partition(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_grid_splits_wavelengths_when_draws_run_out():
    """draws first; one draw on 8 ranks splits the 35 wavelengths (configs 1, 3, 5)."""
    assert sharding.grid_of(4096, 35, 8) == (8, 1)
    assert sharding.grid_of(1, 35, 8) == (1, 8)
    assert sharding.grid_of(3, 35, 8) == (3, 2)
    assert sharding.grid_of(1, 3, 8) == (1, 3)
    for ndraw, nlam, world in ((1, 35, 8), (3, 35, 8), (30, 35, 4), (1, 3, 8), (5, 1, 2)):
        seen = np.zeros((ndraw, nlam), dtype=int)
        for r in range(world):
            (d0, d1), (l0, l1) = sharding.block_of(ndraw, nlam, world, r)
            seen[d0:d1, l0:l1] += 1
        assert (seen == 1).all()


def _inputs(n):
    rng = np.random.default_rng(5)
    return (np.array([500., 700., 900.]), rng.uniform(.4, 2, n), rng.uniform(.3, .95, n), rng.uniform(9, 29, n),
            np.stack([rng.uniform(50, 500, n), rng.uniform(5000, 15000, n)], 1))


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lam, s, g, l0, h = _inputs(n)
    fit, cube, csum = sharding.compute_psf_sharded(lam, s, g, l0, h=h, want_cube=True, compute_fn=fake_compute)
    if rank == 0:
        np.savez(os.path.join(out_dir, 'r0.npz'), fit=fit, cube=cube, csum=csum)
    else:
        assert fit is None and cube is None
        np.savez(os.path.join(out_dir, 'r%d.npz' % rank), csum=csum)
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [1, 5, 8])   # n = 1: the wavelength axis is split instead
def test_two_rank_gather_matches_serial(tmp_path, n):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    lam, s_, g, l0, h = _inputs(n)
    ref_fit, ref_cube = fake_compute(lam, s_, g, l0, h=h)
    got = np.load(tmp_path / 'r0.npz')
    assert np.array_equal(got['fit'], ref_fit)
    assert np.array_equal(got['cube'], ref_cube)
    np.testing.assert_allclose(got['csum'], ref_cube.sum(axis=0), rtol=1e-14)
    np.testing.assert_allclose(np.load(tmp_path / 'r1.npz')['csum'], ref_cube.sum(axis=0), rtol=1e-14)


def test_single_process_passthrough():
    lam, s, g, l0, h = _inputs(3)
    fit, cube, csum = sharding.compute_psf_sharded(lam, s, g, l0, h=h, want_cube=True, compute_fn=fake_compute)
    ref_fit, ref_cube = fake_compute(lam, s, g, l0, h=h)
    assert np.array_equal(fit, ref_fit) and np.array_equal(cube, ref_cube)
    assert np.array_equal(csum, ref_cube.sum(axis=0))
