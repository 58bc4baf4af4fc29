"""Golden-case loader + oracle runner shared by the CPU and GPU parity tests."""
import os

import numpy as np
import torch

from oracle import point_slam_oracle as O
from point_slam_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INTR = synth.TUM_INTRINSICS
CASES = ['mapper_color', 'mapper_geometry', 'tracker_color', 'fixed_radius_zero_depth', 'tum_near_pcl',
         'tum_tracker', 's32_color', 'exposure_tracker', 'exposure_mapper_raw']
# nn_weighting='expo' (decoder.py:154-156): unused by the shipped configs and not differentiable w.r.t. the pose in the reference
# (its in-place masking breaks ExpBackward), so only mapper cases exist (FFMA kernels on the GPU, incl. their weight-gradient phases)
EXPO_CASES = ['expo_mapper', 'expo_mapper_geometry']


def load_scene(dtype=torch.float32, device='cpu'):
    z = np.load(os.path.join(GOLDEN, 'scene.npz'))
    return dict(cloud=torch.from_numpy(z['cloud']).to(device=device, dtype=dtype),
                geo_feats=torch.from_numpy(z['geo_feats'].astype(np.float32)).to(device=device, dtype=dtype),
                col_feats=torch.from_numpy(z['col_feats'].astype(np.float32)).to(device=device, dtype=dtype),
                c2w=z['c2w'])


def load_params(exposure=False, dtype=torch.float32, device='cpu'):
    z = np.load(os.path.join(GOLDEN, 'decoders_exposure.npz' if exposure else 'decoders_base.npz'))
    return {k: torch.from_numpy(z[k]).to(device=device, dtype=dtype) for k in z.files}


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f'case_{name}.npz'))
    c = {k: z[k] for k in z.files}
    for k in ('stage', 'loss_kind'):
        c[k] = str(c[k])
    for k in ('is_tracker', 'use_dynamic_radius', 'encode_rel_pos', 'encode_exposure', 'sample_near_pcl'):
        c[k] = bool(c[k])
    c['S'] = int(c['S'])
    c['nn_weighting'] = str(c['nn_weighting']) if 'nn_weighting' in c else 'distance'
    c['name'] = name
    # ScanNet config uses a wider sampling interval (configs/ScanNet/scannet.yaml:33-34)
    c['near_surface'], c['far_surface'] = (0.96, 1.04) if c['encode_exposure'] else (0.98, 1.02)
    return c


def exposure_mode(c):
    if not c['encode_exposure']:
        return 'none'
    return 'affine' if 'exposure_feat' in c else 'raw'


def run_oracle(c, dtype=torch.float32, tree=None):
    """Forward + loss + backward of one golden case through the oracle.  Returns dict of outputs/grads."""
    scene = load_scene(dtype)
    P = load_params(c['encode_exposure'], dtype)
    P = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    geo = scene['geo_feats'].clone().requires_grad_(True)
    col = scene['col_feats'].clone().requires_grad_(True)
    gt_depth = torch.from_numpy(c['gt_depth']).to(dtype)
    gt_color = torch.from_numpy(c['gt_color']).to(dtype)
    cam = None
    if 'cam_tensor' in c:
        cam = torch.from_numpy(c['cam_tensor']).to(dtype).requires_grad_(True)
        c2w = O.camera_from_tensor(cam)
        i_t = torch.from_numpy(c['pix_i']).to(dtype)
        j_t = torch.from_numpy(c['pix_j']).to(dtype)
        rays_o, rays_d = O.rays_from_uv(i_t, j_t, c2w, INTR['fx'], INTR['fy'], INTR['cx'], INTR['cy'])
    else:
        rays_o = torch.from_numpy(c['rays_o']).to(dtype)
        rays_d = torch.from_numpy(c['rays_d']).to(dtype)
    dyn = torch.from_numpy(c['dynamic_r_query']) if 'dynamic_r_query' in c else None
    ef = None
    if 'exposure_feat' in c:
        ef = torch.from_numpy(c['exposure_feat']).to(dtype).requires_grad_(True)
    zz = torch.from_numpy(c['z_zero_depth']) if 'z_zero_depth' in c else None
    mnn = torch.from_numpy(c['mask_not_near']) if 'mask_not_near' in c else None
    # kNN always on the float32 sample points (the index is float32 in the reference)
    z32, _ = O.sample_z_vals(torch.from_numpy(c['gt_depth']), c['S'], c['near_surface'], c['far_surface'],
                             float(c['near_end']), zz)
    p32 = O.sample_points(rays_o.detach().float(), rays_d.detach().float(), z32)
    dyn_s = None if dyn is None else dyn.reshape(-1, 1).repeat_interleave(c['S'], 0)
    knn = O.find_neighbors(scene['cloud'].float(), p32, float(c['radius_query']), dyn_s, tree=tree)
    depth, var, color, valid, aux = O.render_batch_ray(
        P, rays_d, rays_o, gt_depth, c['stage'], scene['cloud'], geo, col, S=c['S'], is_tracker=c['is_tracker'],
        radius_query=float(c['radius_query']), dynamic_r_query=dyn,
        rand_geo=torch.from_numpy(c['rand_geo']), rand_col=torch.from_numpy(c['rand_col']),
        encode_rel_pos=c['encode_rel_pos'], exposure_mode=exposure_mode(c), exposure_feat=ef,
        sample_near_pcl=c['sample_near_pcl'], near_surface=c['near_surface'], far_surface=c['far_surface'],
        near_end=float(c['near_end']), knn=knn, z_zero_depth=zz, mask_not_near=mnn, return_aux=True,
        weighting=c['nn_weighting'])
    if c['loss_kind'] == 'tracker':
        loss = O.tracker_loss(depth, var, color, gt_depth, gt_color)
    else:
        loss = O.mapper_loss(depth, color, valid, gt_depth, gt_color, c['stage'])
    loss.backward()
    out = dict(depth=depth.detach(), var=var.detach(), color=color.detach(), valid=valid, loss=loss.detach(),
               aux=aux, grad_geo=geo.grad, grad_col=col.grad,
               grad_params={k: v.grad for k, v in P.items() if v.grad is not None})
    if cam is not None:
        out['grad_cam'] = cam.grad
    if ef is not None:
        out['grad_exposure_feat'] = ef.grad
    return out


def rel_err(a, b):
    """max |a-b| / max(|b|) -- the 'relative to the tensor scale' error the 1e-4 north-star bar is read with."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
