"""NIfTI-1 / legacy-VTK sample writers (reference logger/logger.py:35-102,215-238)"""
import gzip
import struct

import numpy as np
import torch


def test_nifti_round_trip_and_header(tmp_path):
    from irsgmcmc_b200.logger import load_im_from_disk, save_im_to_disk
    rng = np.random.default_rng(0)
    for arr in (rng.random((5, 6, 7), dtype=np.float32), rng.integers(0, 60, (4, 3, 2)).astype(np.int16),
                rng.random((3, 3, 3)) > 0.5):
        for name in ('a.nii', 'a.nii.gz'):
            p = tmp_path / name
            save_im_to_disk(torch.from_numpy(arr), str(p), spacing=torch.tensor([1.5, 2.0, 2.5]))
            back, sp = load_im_from_disk(str(p))
            assert back.shape == arr.shape and np.array_equal(back, arr.astype(back.dtype)) and np.allclose(sp, (1.5, 2.0, 2.5))
    raw = gzip.open(tmp_path / 'a.nii.gz', 'rb').read()
    assert struct.unpack_from('<i', raw, 0)[0] == 348 and raw[344:348] == b'n+1\0' and raw[123] == 2   # units: mm
    assert struct.unpack_from('<h', raw, 254)[0] == 2 and struct.unpack_from('<f', raw, 108)[0] == 352.0
    # data are in NIfTI (first index fastest) order
    x = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    save_im_to_disk(x, str(tmp_path / 'o.nii'))
    raw = open(tmp_path / 'o.nii', 'rb').read()
    assert np.array_equal(np.frombuffer(raw, '<f4', 3, 352), [x[0, 0, 0], x[1, 0, 0], x[0, 1, 0]])


def test_vtk_field_round_trip_and_sample_names(tmp_path):
    from irsgmcmc_b200.logger import load_field_from_disk, save_field_to_disk, save_grid_to_disk, save_sample
    f = np.random.default_rng(1).standard_normal((3, 4, 5, 6)).astype(np.float32)
    p = tmp_path / 'f.vtk'
    save_field_to_disk(torch.from_numpy(f), str(p), spacing=(1.0, 2.0, 3.0))
    back, sp = load_field_from_disk(str(p))
    assert np.array_equal(back, f) and sp == [1.0, 2.0, 3.0]
    head = open(p, 'rb').read(200).decode('ascii', 'ignore')
    assert head.startswith('# vtk DataFile Version 3.0') and 'DATASET STRUCTURED_POINTS' in head and 'DIMENSIONS 4 5 6' in head
    save_grid_to_disk(torch.from_numpy(f), str(tmp_path / 'g.vtk'))
    assert b'DATASET STRUCTURED_GRID' in open(tmp_path / 'g.vtk', 'rb').read(120)
    paths = save_sample(str(tmp_path / 's'), (2.0, 2.0, 2.0), 12, torch.rand(1, 1, 4, 5, 6), torch.from_numpy(f)[None],
                        torch.rand(1, 4, 5, 6), model='MCMC', chain_no=3)
    assert paths['displacement'].endswith('chain_3_sample_0000012_displacement.vtk')
    d, _ = load_field_from_disk(paths['displacement'])
    assert np.allclose(d, 2.0 * f)      # scaled by spacing[0] like the reference (logger/logger.py:223)


def test_statistics_and_posterior_mean_files(tmp_path):
    """the reference's file names for the sample statistics and the posterior-mean registration (logger/logger.py:110-131,198-208)"""
    from irsgmcmc_b200.logger import (load_field_from_disk, load_im_from_disk, save_displacement_mean_and_std_dev,
                                      save_variational_posterior_mean)
    rng = np.random.default_rng(2)
    mean, std = (rng.standard_normal((3, 4, 5, 6)).astype(np.float32) for _ in range(2))
    mask = torch.from_numpy(rng.random((1, 1, 4, 5, 6)) > 0.5)
    paths = save_displacement_mean_and_std_dev(str(tmp_path), (2.0, 2.0, 2.0), torch.from_numpy(mean), torch.from_numpy(std), mask, 'VI')
    assert sorted(p.name for p in tmp_path.iterdir()) == ['VI_sample_mean.vtk', 'VI_sample_mean_masked.vtk', 'VI_sample_std_dev.vtk',
                                                           'VI_sample_std_dev_masked.vtk']
    back, sp = load_field_from_disk(paths['mean'])
    assert np.allclose(back, 2.0 * mean) and sp == [2.0, 2.0, 2.0]
    back, _ = load_field_from_disk(paths['std_dev_masked'])
    assert np.allclose(back, 2.0 * std * mask[0].numpy())
    assert set(save_displacement_mean_and_std_dev(str(tmp_path / 'nomask'), (1, 1, 1), mean, std, None, 'MCMC')) == {'mean', 'std_dev'}
    im, d = rng.random((1, 1, 4, 5, 6)).astype(np.float32), rng.standard_normal((1, 3, 4, 5, 6)).astype(np.float32)
    paths = save_variational_posterior_mean(str(tmp_path / 'mu'), torch.tensor([1.5, 1.5, 1.5]), torch.from_numpy(im), torch.from_numpy(d))
    assert paths['im_moving_warped_mu'].endswith('im_moving_warped_mu.nii.gz') and paths['displacement_mu'].endswith('displacement_mu.vtk')
    back, _ = load_im_from_disk(paths['im_moving_warped_mu'])
    assert np.array_equal(back, im[0, 0])
    back, _ = load_field_from_disk(paths['displacement_mu'])
    assert np.allclose(back, 1.5 * d[0])


def test_average_surface_distance_against_brute_force():
    """calc_ASD_host (the reference's LabelContour + average Hausdorff distance, utils/util.py:171-176, restated with scipy) against a
    brute-force evaluation of the same definition; parity with SimpleITK itself is unpinned (absent from the image)"""
    from irsgmcmc_b200.utils.util import calc_ASD_host, _label_contour
    rng = np.random.default_rng(3)
    n = 14
    zz, yy, xx = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing='ij')
    a = (zz - 6.0) ** 2 + (yy - 7.0) ** 2 + (xx - 6.5) ** 2 <= 16.0
    b = (zz - 7.0) ** 2 / 1.5 + (yy - 6.0) ** 2 + (xx - 7.5) ** 2 <= 14.0
    spacing = (1.0, 1.5, 2.0)          # (x, y, z) like the reference's im_spacing
    ca, cb = _label_contour(a), _label_contour(b)
    assert ca.sum() < a.sum() and a[6, 7, 6] and not ca[6, 7, 6]      # the interior is not contour
    pa = np.argwhere(ca) * np.array(spacing[::-1])
    pb = np.argwhere(cb) * np.array(spacing[::-1])
    dist = np.sqrt(((pa[:, None, :] - pb[None, :, :]) ** 2).sum(-1))
    want = 0.5 * (dist.min(1).mean() + dist.min(0).mean())
    assert abs(calc_ASD_host(a, b, spacing) - want) < 1e-9
    assert calc_ASD_host(a, a, spacing) == 0.0
    assert calc_ASD_host(a, np.zeros_like(a), spacing) == float('inf')
