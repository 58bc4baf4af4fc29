"""Pins oracle/point_slam_oracle.py against vectors frozen from the UNMODIFIED reference
(tests/golden/*, written by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import point_slam_oracle as O
from tests import cases as C

TOL = 2e-5      # oracle and reference are both fp32 torch on the CPU: only op-fusion order differs


@pytest.mark.parametrize('name', C.CASES + C.EXPO_CASES)
def test_case_forward_and_grads(name):
    c = C.load_case(name)
    o = C.run_oracle(c)
    assert np.array_equal(o['valid'].numpy(), c['valid'])
    assert C.rel_err(o['depth'], c['depth']) < TOL
    assert C.rel_err(o['var'], c['var']) < 5e-4           # var = sum w (z-depth)^2: cancellation-amplified
    assert C.rel_err(o['color'], c['color']) < TOL
    assert abs(float(o['loss']) - float(c['loss'])) / abs(float(c['loss'])) < TOL
    if 'grad.cam_tensor' in c:
        assert C.rel_err(o['grad_cam'], c['grad.cam_tensor']) < 2e-4
    if 'grad.exposure_feat' in c:
        assert C.rel_err(o['grad_exposure_feat'], c['grad.exposure_feat']) < 1e-4
    for key, g in (('geo', o['grad_geo']), ('col', o['grad_col'])):
        if f'grad.{key}_rows' in c:
            rows, vals = c[f'grad.{key}_rows'], c[f'grad.{key}_vals']
            dense = np.zeros(tuple(g.shape), np.float32)
            dense[rows] = vals
            assert C.rel_err(g, dense) < 1e-4, key
    n_checked = 0
    for k, g in o['grad_params'].items():
        gk = 'grad.' + k
        if gk in c:
            assert C.rel_err(g, c[gk]) < 2e-4, k
            n_checked += 1
    assert n_checked >= (10 if c['stage'] == 'color' or name in C.CASES else 5)


def test_aux_knn_contract():
    z = np.load(C.GOLDEN + '/aux.npz')
    scene = C.load_scene()
    D, I, n = O.find_neighbors(scene['cloud'], torch.from_numpy(z['knn_pts']), 0.08)
    assert np.array_equal(I.numpy(), z['knn_I'])
    assert np.array_equal(D.numpy(), z['knn_D'])
    assert np.array_equal(n.numpy(), z['knn_n'])
    # brute force and kd-tree candidate paths agree bit for bit
    from scipy.spatial import cKDTree
    tree = cKDTree(scene['cloud'].double().numpy())
    D2, I2 = O.knn_exact(scene['cloud'], torch.from_numpy(z['knn_pts']), tree=tree)
    assert np.array_equal(I2.numpy(), I.numpy()) and np.array_equal(D2.numpy(), D.numpy())


def test_aux_rays_and_composite():
    z = np.load(C.GOLDEN + '/aux.npz')
    scene = C.load_scene()
    c2w = torch.tensor(scene['c2w'][:3, :4], dtype=torch.float32)
    ro, rd = O.rays_full_image(12, 16, 20.0, 21.0, 7.5, 5.5, c2w)
    assert np.array_equal(rd.numpy(), z['get_rays_d']) and np.array_equal(ro.numpy(), z['get_rays_o'])
    d, v, rgb, w = O.composite(torch.from_numpy(z['comp_raw']), torch.from_numpy(z['comp_z']))
    for a, b in ((d, 'comp_depth'), (v, 'comp_var'), (rgb, 'comp_rgb'), (w, 'comp_w')):
        assert C.rel_err(a, z[b]) < 1e-6


def test_aux_add_points_and_near_pcl():
    z = np.load(C.GOLDEN + '/aux.npz')
    scene = C.load_scene()
    ro, rd, gd = (torch.from_numpy(z[k]) for k in ('add_rays_o', 'add_rays_d', 'add_depth'))
    r_add = torch.from_numpy(z['add_r_add'])
    keep1, new1 = O.add_points(scene['cloud'], ro, rd, gd, dynamic_radius=r_add)
    assert int(keep1.sum()) == int(z['add_kept1'])
    assert np.array_equal(new1.numpy(), z['add_new1'])
    cloud2 = torch.cat([scene['cloud'], new1], 0)
    keep2, new2 = O.add_points(cloud2, ro, rd, gd, is_pts_grad=True)
    assert int(keep2.sum()) == int(z['add_kept2'])
    assert np.array_equal(new2.numpy(), z['add_new2'])
    zz, inv = O.sample_near_pcl(scene['cloud'], ro[:96], rd[:96], 0.3, 4.2, 5)
    assert np.array_equal(inv.numpy(), z['snp_invalid'])
    assert np.allclose(zz.numpy(), z['snp_z'], rtol=0, atol=1e-6)


def test_fp64_noise_floor():
    """fp32 vs fp64 evaluation of the same algorithm: documents how much of the 1e-4 budget rounding eats."""
    c = C.load_case('mapper_color')
    o32 = C.run_oracle(c, torch.float32)
    o64 = C.run_oracle(c, torch.float64)
    assert C.rel_err(o32['depth'], o64['depth']) < 1e-5
    assert C.rel_err(o32['color'], o64['color']) < 1e-4


def test_frustum_selection_matches_reference():
    """oracle.frustum_indices == Mapper.get_mask_from_c2w of the unmodified reference (cv2.remap included), bit for bit."""
    z = np.load(C.GOLDEN + '/frustum.npz')
    H, W, fx, fy, cx, cy = z['intr']
    for k in range(2):
        idx = O.frustum_indices(z['cloud'], z[f'c2w{k}'], z[f'depth{k}'], int(H), int(W), fx, fy, cx, cy, edge=int(z[f'edge{k}']))
        assert np.array_equal(idx, z[f'indices{k}']), k
        assert 500 < idx.size < z['cloud'].shape[0] // 2
