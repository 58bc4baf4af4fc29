"""Drop-in for the hot-path parts of the reference `src/common.py`: ray generation, pixel sampling, pose
conversions and the alpha composite.  Ray/pose helpers are a handful of torch ops on (R,3) tensors (differentiable
w.r.t. the pose, common.py:40-56, :225-267); the composite runs in the CUDA library.
Host-side gradient-based pixel pickers (common.py:92-159, :186-222) are out of scope.
"""
import random

import numpy as np
import torch

from .. import ops


def setup_seed(seed):
    """common.py:10-16"""
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def _camera_dirs(i, j, fx, fy, cx, cy):
    return torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)


def get_rays_from_uv(i, j, c2w, fx, fy, cx, cy, device):
    """common.py:40-56 -- rays for flattened pixel coordinates (i = column, j = row); rays_d is NOT normalised."""
    if isinstance(c2w, np.ndarray):
        c2w = torch.from_numpy(c2w).to(device)
    dirs = _camera_dirs(i, j, fx, fy, cx, cy).to(device).reshape(-1, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def get_rays(H, W, fx, fy, cx, cy, c2w, device, crop_edge=0):
    """common.py:339-356 -- rays for a whole image, (H, W, 3)."""
    if isinstance(c2w, np.ndarray):
        c2w = torch.from_numpy(c2w)
    i, j = torch.meshgrid(torch.linspace(crop_edge, W - 1 - crop_edge, W - 2 * crop_edge),
                          torch.linspace(crop_edge, H - 1 - crop_edge, H - 2 * crop_edge), indexing='ij')
    dirs = _camera_dirs(i.t(), j.t(), fx, fy, cx, cy).to(device).reshape(H - 2 * crop_edge, W - 2 * crop_edge, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3].to(device), -1)
    rays_o = c2w[:3, -1].to(device).expand(rays_d.shape)
    return rays_o, rays_d


def select_uv(i, j, n, depth, color, device='cuda:0'):
    """common.py:59-74 -- n uniformly random pixels (device RNG, same draw as the reference)."""
    i, j = i.reshape(-1), j.reshape(-1)
    indices = torch.randint(i.shape[0], (n,), device=device).clamp(0, i.shape[0])
    return i[indices], j[indices], depth.reshape(-1)[indices], color.reshape(-1, 3)[indices]


def get_sample_uv(H0, H1, W0, W1, n, depth, color, device='cuda:0'):
    """common.py:77-89"""
    depth, color = depth[H0:H1, W0:W1], color[H0:H1, W0:W1]
    i, j = torch.meshgrid(torch.linspace(W0, W1 - 1, W1 - W0).to(device), torch.linspace(H0, H1 - 1, H1 - H0).to(device),
                          indexing='ij')
    return select_uv(i.t(), j.t(), n, depth, color, device=device)


def get_samples(H0, H1, W0, W1, n, fx, fy, cx, cy, c2w, depth, color, device, depth_filter=False, return_index=False,
                depth_limit=None):
    """common.py:162-183"""
    i, j, sample_depth, sample_color = get_sample_uv(H0, H1, W0, W1, n, depth, color, device=device)
    rays_o, rays_d = get_rays_from_uv(i, j, c2w, fx, fy, cx, cy, device)
    if depth_filter:
        mask = sample_depth > 0
        if depth_limit is not None:
            mask = mask & (sample_depth < depth_limit)
        rays_o, rays_d, sample_depth, sample_color = rays_o[mask], rays_d[mask], sample_depth[mask], sample_color[mask]
        i, j = i[mask], j[mask]
    if return_index:
        return rays_o, rays_d, sample_depth, sample_color, i.to(torch.int64), j.to(torch.int64)
    return rays_o, rays_d, sample_depth, sample_color


def quad2rotation(quad):
    """common.py:225-248 -- batch quaternion (w,x,y,z) -> rotation, differentiable."""
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    m = [1 - two_s * (qj ** 2 + qk ** 2), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
         two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr),
         two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2)]
    return torch.stack(m, -1).reshape(-1, 3, 3)


def get_camera_from_tensor(inputs):
    """common.py:251-267 -- [quat, T] -> 3x4 (or N x 3x4)."""
    single = len(inputs.shape) == 1
    x = inputs.unsqueeze(0) if single else inputs
    RT = torch.cat([quad2rotation(x[:, :4]), x[:, 4:, None]], 2)
    return RT[0] if single else RT


def get_tensor_from_camera(RT, Tquad=False):
    """common.py:270-295"""
    from scipy.spatial.transform import Rotation
    dev = None
    if isinstance(RT, torch.Tensor):
        dev = RT.device if RT.is_cuda else None
        RT = RT.detach().cpu().numpy()
    quad = np.roll(Rotation.from_matrix(RT[:3, :3]).as_quat(), 1)
    T = RT[:3, 3]
    out = torch.from_numpy(np.concatenate([T, quad] if Tquad else [quad, T], 0)).float()
    return out.to(dev) if dev is not None else out


def raw2outputs_nerf_color(raw, z_vals, rays_d, device='cuda:0', coef=0.1):
    """common.py:298-336 -> depth_map, depth_var, rgb_map, weights.  (rays_d only feeds dead code there, :316-321.)
    Unlike the reference this does not overwrite raw[..., -1] in place; no caller reads it afterwards."""
    return ops.composite(raw, z_vals, None, coef)
