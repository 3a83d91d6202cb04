"""Multi-GPU parity of sharding.compute_psf_sharded on real devices (NCCL): the sharded result must
equal the single-GPU result - bit for bit when draws are split, to rounding-noise level when the
wavelength axis is split (a wavelength's FP32 pairing partner can change).
Skipped on boxes with fewer than 2 GPUs; the host logic has world-size-2 gloo tests in test_sharding.py."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs():
    rng = np.random.default_rng(77)
    n = 7
    return (np.linspace(490, 930, 6), rng.uniform(.4, 2, n), rng.uniform(.3, .95, n), rng.uniform(9, 29, n),
            np.stack([rng.uniform(50, 500, n), rng.uniform(5000, 15000, n)], 1))


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from muse_psfr_b200 import psfrec, sharding
    psfrec.set_device(rank)
    lam, s, g, l0, h = _inputs()
    fit, cube, csum = sharding.compute_psf_sharded(lam, s, g, l0, h=h, want_cube=True)          # draws split
    fit1, cube1, _ = sharding.compute_psf_sharded(lam, s[:1], g[:1], l0[:1], h=h[:1], want_cube=True)   # wavelengths split
    fit_only, none_cube, none_sum = sharding.compute_psf_sharded(lam, s, g, l0, h=h, want_sum=False)
    if rank == 0:
        assert none_cube is None and none_sum is None
        np.savez(os.path.join(out_dir, 'r0.npz'), fit=fit, cube=cube, csum=csum, fit1=fit1, cube1=cube1, fit_only=fit_only)
    else:
        assert fit is None and cube is None and fit1 is None and fit_only is None
    dist.destroy_process_group()


def test_sharded_equals_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from muse_psfr_b200 import _lib, psfrec
    psfrec.set_device(0)
    lam, s_, g, l0, h = _inputs()
    ref_fit, ref_cube = psfrec.compute_psf_batch(lam, s_, g, l0, h=h)
    got = np.load(tmp_path / 'r0.npz')
    assert np.array_equal(got['fit'], ref_fit)          # bit for bit: draws are independent of their batch
    assert np.array_equal(got['cube'], ref_cube)
    assert np.array_equal(got['fit_only'], ref_fit)
    np.testing.assert_allclose(got['csum'], ref_cube.sum(axis=0), rtol=1e-13)
    np.testing.assert_allclose(got['cube1'][0], ref_cube[0], rtol=0, atol=1e-13 * ref_cube[0].max())
    np.testing.assert_allclose(got['fit1'][0, :, _lib.FIT_FWHM], ref_fit[0, :, _lib.FIT_FWHM], rtol=1e-9)
    psfrec.release_contexts()
