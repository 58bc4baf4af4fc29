"""CPU oracle for the Point-SLAM render hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional (stateless) restatement, in plain PyTorch on the CPU,
of the algorithm the reference runs for one `render_batch_ray` call:

    ray sampling -> radius-kNN (k=8) -> IDW interpolation (+ per-neighbour colour
    MLP) -> geometry / colour MLP decode -> alpha composite  (+ autograd backward)

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  The product path (`point_slam_b200/`) never
does: it fails loudly when the CUDA library is missing.

Parity status: **pinned against the reference itself**.  `oracle/make_golden.py`
imports the unmodified reference from /root/reference (with `faiss`/`skimage`
import stubs and an exact-kNN shim for the un-vendored faiss-gpu 1.7.2
dependency, env.yaml:96) and freezes its outputs under `tests/golden/`;
`tests/test_oracle_vs_golden.py` checks this restatement against those files.
The kNN contract is *exact* search (what faiss IVF approximates)
with the canonical fp32 distance  D = fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz))
and ties broken towards the smaller point index.

Every function cites the reference lines it follows (paths relative to the
reference root).  All arithmetic is done in the dtype of the inputs (float32 for
parity, float64 to bound fp32 noise); parameters are passed as a flat dict whose
keys are the reference `state_dict` names (`geo_decoder.*`, `color_decoder.*`)
plus `color_decoder.embedder._B`, which the reference keeps as a plain tensor
(src/conv_onet/models/decoder.py:27-28).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

NN_NUM = 8          # configs/point_slam.yaml:107
MIN_NN_NUM = 2      # configs/point_slam.yaml:108
C_DIM = 32          # configs/point_slam.yaml:10


# --------------------------------------------------------------------------- #
# rays                                                                        #
# --------------------------------------------------------------------------- #
def rays_from_uv(i, j, c2w, fx, fy, cx, cy):
    """src/common.py:40-56 -- un-normalised world-space ray per pixel (i=col, j=row)."""
    cam = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    rays_d = (cam.reshape(-1, 1, 3) * c2w[:3, :3]).sum(-1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def rays_full_image(H, W, fx, fy, cx, cy, c2w, crop_edge=0):
    """src/common.py:339-356 -- rays for every pixel, row-major (H, W, 3)."""
    ii, jj = torch.meshgrid(torch.linspace(crop_edge, W - 1 - crop_edge, W - 2 * crop_edge),
                            torch.linspace(crop_edge, H - 1 - crop_edge, H - 2 * crop_edge), indexing='ij')
    ii, jj = ii.t().to(c2w.dtype), jj.t().to(c2w.dtype)
    cam = torch.stack([(ii - cx) / fx, -(jj - cy) / fy, -torch.ones_like(ii)], -1)
    rays_d = (cam.reshape(H - 2 * crop_edge, W - 2 * crop_edge, 1, 3) * c2w[:3, :3]).sum(-1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def quad_to_rotation(quad):
    """src/common.py:225-248 (device line :238 dropped so it also runs on the CPU)."""
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    rows = [
        1 - two_s * (qj ** 2 + qk ** 2), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
        two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr),
        two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2),
    ]
    return torch.stack(rows, -1).reshape(-1, 3, 3)


def camera_from_tensor(cam7):
    """src/common.py:251-267 -- [quat(4), T(3)] -> 3x4 [R|T]."""
    single = cam7.dim() == 1
    x = cam7.unsqueeze(0) if single else cam7
    RT = torch.cat([quad_to_rotation(x[:, :4]), x[:, 4:, None]], 2)
    return RT[0] if single else RT


# --------------------------------------------------------------------------- #
# ray marching (sample placement)                                             #
# --------------------------------------------------------------------------- #
def surface_t_vals(S, dtype=torch.float32):
    """t = linspace(0,1,S) exactly as torch evaluates it on the CPU (Renderer.py:137-138)."""
    return torch.linspace(0.0, 1.0, steps=S, dtype=torch.float32).to(dtype)


def sample_z_vals(gt_depth, S, near_surface=0.98, far_surface=1.02, near_end=0.3,
                  z_zero_depth: Optional[torch.Tensor] = None):
    """src/utils/Renderer.py:108-170.

    gt_depth (R,) -> z_vals (R,S), gt_non_zero_mask (R,).  Rays with depth>0 get
    z = near*D*(1-t) + far*D*t in exactly that operator order (:140-142); rays with
    depth<=0 get `z_zero_depth` rows when given (the `sample_near_pcl` result,
    :154-164) else linspace(near_end, max(far), S) (:165-168) with
    far = min(5*mean(D), max(1.2*D)) (:111-112).
    """
    R = gt_depth.shape[0]
    D = gt_depth.reshape(-1, 1)
    nz = (D > 0).squeeze(-1)
    t = surface_t_vals(S, gt_depth.dtype)
    z = torch.zeros(R, S, dtype=gt_depth.dtype)
    Dn = D[nz].repeat(1, S)
    z[nz] = near_surface * Dn * (1. - t) + far_surface * Dn * t
    if int(nz.sum()) < R:
        if z_zero_depth is not None:
            z[~nz] = z_zero_depth.to(z.dtype)
        else:
            far = torch.minimum(5 * gt_depth.mean(), torch.max(gt_depth * 1.2)).float()
            z[~nz] = torch.linspace(near_end, float(far), steps=S).to(z.dtype).repeat(int((~nz).sum()), 1)
    return z, nz


def sample_points(rays_o, rays_d, z_vals):
    """src/utils/Renderer.py:172-174 -- p = o + d*z, flattened ray-major (R*S,3)."""
    return (rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]).reshape(-1, 3)


# --------------------------------------------------------------------------- #
# exact radius-kNN (stands in for faiss GpuIndexIVFFlat.search)               #
# --------------------------------------------------------------------------- #
def canonical_sqdist(cloud: np.ndarray, q: np.ndarray) -> np.ndarray:
    """fp32, no FMA: ((dx*dx + dy*dy) + dz*dz) with dx = c - q.  cloud (N,3), q (B,3) -> (B,N)."""
    c = cloud.astype(np.float32)
    q = q.astype(np.float32)
    dx = c[None, :, 0] - q[:, None, 0]
    dy = c[None, :, 1] - q[:, None, 1]
    dz = c[None, :, 2] - q[:, None, 2]
    return (dx * dx + dy * dy) + dz * dz


def knn_exact(cloud_pos, pos, k=NN_NUM, chunk=2048, tree=None):
    """Exact k-NN with the contract of src/neural_point.py:189-197 (`index.search`).

    Returns D (M,k) float32 squared-L2 ascending, I (M,k) int64.  Ties (equal fp32 D)
    go to the smaller index.  Slots that do not exist (N<k) carry I=-1, D=FLT_MAX
    (the faiss convention).  For large clouds pass `tree` (scipy cKDTree over
    float64 positions): k+24 candidates are fetched, re-scored with the canonical
    fp32 formula and the margin is asserted to be sufficient.
    """
    cloud = np.ascontiguousarray(cloud_pos.detach().cpu().numpy() if torch.is_tensor(cloud_pos) else cloud_pos,
                                 dtype=np.float32)
    q_all = np.ascontiguousarray(pos.detach().cpu().numpy() if torch.is_tensor(pos) else pos, dtype=np.float32)
    M, N = q_all.shape[0], cloud.shape[0]
    D_out = np.full((M, k), np.finfo(np.float32).max, dtype=np.float32)
    I_out = np.full((M, k), -1, dtype=np.int64)
    if N == 0 or M == 0:
        return torch.from_numpy(D_out), torch.from_numpy(I_out)
    kk = min(k, N)
    if tree is not None and N > k + 24:
        kc = k + 24
        _, cand = tree.query(q_all.astype(np.float64), k=kc, workers=-1)
        cand = np.sort(cand, axis=1)                      # ascending index -> stable sort breaks ties right
        c = cloud[cand]                                   # (M,kc,3)
        dx = c[..., 0] - q_all[:, None, 0]
        dy = c[..., 1] - q_all[:, None, 1]
        dz = c[..., 2] - q_all[:, None, 2]
        Dc = (dx * dx + dy * dy) + dz * dz
        order = np.argsort(Dc, axis=1, kind='stable')
        Ds = np.take_along_axis(Dc, order, 1)
        Is = np.take_along_axis(cand, order, 1)
        # the farthest candidate (by the tree's float64 metric) must not be able to beat slot k-1
        assert np.all(Ds[:, kc - 1] > Ds[:, k - 1]), 'kNN margin too small for canonical re-ranking'
        D_out[:, :kk], I_out[:, :kk] = Ds[:, :kk], Is[:, :kk]
        return torch.from_numpy(D_out), torch.from_numpy(I_out)
    for s in range(0, M, chunk):
        q = q_all[s:s + chunk]
        Dm = canonical_sqdist(cloud, q)
        if N > 4 * kk:
            part = np.argpartition(Dm, kk + min(16, N - kk - 1), axis=1)[:, :kk + min(16, N - kk - 1) + 1]
            part = np.sort(part, axis=1)
            Dp = np.take_along_axis(Dm, part, 1)
            order = np.argsort(Dp, axis=1, kind='stable')
            Ds = np.take_along_axis(Dp, order, 1)
            Is = np.take_along_axis(part, order, 1)
            # ties with an un-fetched element would break the index tie rule: verify none
            kth = Ds[:, kk - 1:kk]
            assert np.all((Dm <= kth).sum(1) <= part.shape[1]), 'tie overflow in knn_exact'
        else:
            order = np.argsort(Dm, axis=1, kind='stable')
            Ds = np.take_along_axis(Dm, order, 1)
            Is = order
        D_out[s:s + chunk, :kk], I_out[s:s + chunk, :kk] = Ds[:, :kk], Is[:, :kk]
    return torch.from_numpy(D_out), torch.from_numpy(I_out)


def radius_sq_f64(M, radius=None, dynamic_radius=None):
    """Squared radius exactly as the comparison sees it (neural_point.py:208-213):
    a float64 per-query tensor when `dynamic_radius` is given (compared in float64),
    else the Python scalar radius**2 rounded to float32 (torch casts a Python scalar
    to the tensor dtype before comparing)."""
    if dynamic_radius is not None:
        return dynamic_radius.reshape(-1).double() ** 2
    return torch.full((M,), float(np.float32(radius ** 2)), dtype=torch.float64)


def find_neighbors(cloud_pos, pos, radius=0.08, dynamic_radius=None, k=NN_NUM, tree=None):
    """src/neural_point.py:169-215 -> (D, I, neighbor_num) with neighbor_num = #(D < r^2) strict."""
    D, I = knn_exact(cloud_pos, pos, k, tree=tree)
    r2 = radius_sq_f64(D.shape[0], radius, dynamic_radius)
    neighbor_num = (D.double() < r2[:, None]).sum(-1).int()
    return D, I, neighbor_num


# --------------------------------------------------------------------------- #
# decoder                                                                     #
# --------------------------------------------------------------------------- #
def fourier(x, B, with_cos):
    """src/conv_onet/models/decoder.py:30-37."""
    y = (2 * math.pi * x) @ B
    return torch.cat((torch.sin(y), torch.cos(y)), -1) if with_cos else torch.sin(y)


def softplus100(x):
    """nn.Softplus(beta=100) with PyTorch's threshold=20 linearisation (decoder.py:231,335)."""
    return F.softplus(x, beta=100)


def idw_weights(D, r2, weighting='distance'):
    """decoder.py:152-160 -- 1/(D+1e-10) (or exp(-20 sqrt(D)) for nn_weighting='expo', :154-156), zero where D > r^2
    (strict, float64 compare when r^2 is float64), then L1-normalise with eps 1e-12."""
    w = 1.0 / (D + 1e-10) if weighting == 'distance' else torch.exp(-20 * torch.sqrt(D))
    w = torch.where(D.double() > r2.reshape(-1, 1), torch.zeros_like(w), w)
    return F.normalize(w, p=1, dim=1)


def _lin(P: Params, name: str, x):
    return F.linear(x, P[name + '.weight'], P[name + '.bias'])


def interp_geo(P, p, D_knn, I, neighbor_num, r2, geo_feats, cloud_pos, is_tracker, rand_vec, weighting='distance'):
    """MLP_geometry.get_feature_at_pos, decoder.py:130-173."""
    D = D_knn.to(p.dtype)
    if is_tracker:                                             # :143-148 (pose gradient path)
        D = torch.square(cloud_pos[I] - p.reshape(-1, 1, 3)).sum(-1)
    has_neighbors = neighbor_num > MIN_NN_NUM - 1              # :150
    w = idw_weights(D, r2, weighting).unsqueeze(-1)
    c = (w * geo_feats[I]).sum(1)
    c = torch.where(has_neighbors[:, None], c, rand_vec.to(c.dtype)[None, :])   # :170-171
    return c, has_neighbors


def interp_col(P, p, D_knn, I, neighbor_num, r2, col_feats, cloud_pos, is_tracker, rand_vec,
               encode_rel_pos=True, weighting='distance'):
    """MLP_color.get_feature_at_pos, decoder.py:341-390."""
    D = D_knn.to(p.dtype)
    if is_tracker:
        D = torch.square(cloud_pos[I] - p.reshape(-1, 1, 3)).sum(-1)
    has_neighbors = neighbor_num > MIN_NN_NUM - 1
    w = idw_weights(D, r2, weighting).unsqueeze(-1)
    f = col_feats[I]
    if encode_rel_pos:                                         # :373-381
        rel = cloud_pos[I] - p[:, None, :]
        emb = fourier(rel.reshape(-1, 3), P['color_decoder.embedder_rel_pos._B'], True)
        x = torch.cat([emb.reshape(rel.shape[0], -1, emb.shape[-1]), f], -1)
        h = softplus100(_lin(P, 'color_decoder.mlp_col_neighbor.linear1', x))
        f = _lin(P, 'color_decoder.mlp_col_neighbor.linear2', h)
    c = (w * f).sum(1)
    c = torch.where(has_neighbors[:, None], c, rand_vec.to(c.dtype)[None, :])   # :387-388
    return c, has_neighbors


def geo_trunk(P, p, c):
    """MLP_geometry.forward trunk, decoder.py:203-221 (ReLU, skip-cat after block 2)."""
    e = fourier(p, P['geo_decoder.embedder._B'], False)
    h = e
    for i in range(5):
        h = F.relu(_lin(P, f'geo_decoder.pts_linears.{i}', h))
        h = h + _lin(P, f'geo_decoder.fc_c.{i}', c)
        if i == 2:
            h = torch.cat([e, h], -1)
    return _lin(P, 'geo_decoder.output_linear', h).squeeze(-1)


def col_trunk(P, p, c, exposure_mode='none', exposure_feat=None):
    """MLP_color.forward trunk, decoder.py:411-449 (Softplus beta=100).
    exposure_mode: 'none' -> sigmoid(out) (:447); 'affine' -> sigmoid(out@rot+trans) (:433-438);
    'raw' -> out (encode_exposure with exposure_feat None, :439-445)."""
    e = fourier(p, P['color_decoder.embedder._B'], True)
    h = e
    for i in range(5):
        h = softplus100(_lin(P, f'color_decoder.pts_linears.{i}', h))
        h = h + _lin(P, f'color_decoder.fc_c.{i}', c)
        if i == 2:
            h = torch.cat([e, h], -1)
    out = _lin(P, 'color_decoder.output_linear', h)
    if exposure_mode == 'none':
        return torch.sigmoid(out)
    if exposure_mode == 'raw':
        return out
    a = _lin(P, 'color_decoder.mlp_exposure.linear2',
             softplus100(_lin(P, 'color_decoder.mlp_exposure.linear1', exposure_feat)))
    a = a.reshape(-1)
    return torch.sigmoid(out @ a[:9].reshape(3, 3) + a[-3:])


def point_forward(P, p, stage, knn, r2, geo_feats, col_feats, cloud_pos, S, is_tracker,
                  rand_geo, rand_col, encode_rel_pos=True, exposure_mode='none', exposure_feat=None, weighting='distance'):
    """POINT.forward, decoder.py:476-518 -> raw (M,4), ray_mask (R,), point_mask (M,)."""
    D, I, nnum = knn
    c_g, has_nb = interp_geo(P, p, D, I, nnum, r2, geo_feats, cloud_pos, is_tracker, rand_geo, weighting)
    occ = geo_trunk(P, p, c_g)
    ray_mask = ~(has_nb.view(-1, S).sum(1) < int(S / 2 + 1))           # :200-201
    if stage == 'geometry':
        raw = torch.cat([torch.zeros(occ.shape[0], 3, dtype=occ.dtype), occ[:, None]], -1)
        return raw, ray_mask, has_nb
    c_c, _ = interp_col(P, p, D, I, nnum, r2, col_feats, cloud_pos, is_tracker, rand_col, encode_rel_pos, weighting)
    rgb = col_trunk(P, p, c_c, exposure_mode, exposure_feat)
    return torch.cat([rgb, occ[:, None]], -1), ray_mask, has_nb


# --------------------------------------------------------------------------- #
# composite                                                                   #
# --------------------------------------------------------------------------- #
def composite(raw, z_vals, coef=0.1):
    """raw2outputs_nerf_color, src/common.py:298-336 (the `dists` lines :316-321 are dead code)."""
    rgb = raw[..., :3]
    alpha = torch.sigmoid(coef * raw[..., 3])
    ones = torch.ones(alpha.shape[0], 1, dtype=alpha.dtype)
    trans = torch.cumprod(torch.cat([ones, 1. - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * trans
    wsum = w.sum(-1, keepdim=True) + 1e-10
    rgb_map = (w[..., None] * rgb).sum(-2) / wsum
    depth = (w * z_vals).sum(-1) / wsum.squeeze(-1)
    tmp = z_vals - depth.unsqueeze(-1)
    var = (w * tmp * tmp).sum(1)
    return depth, var, rgb_map, w


def mask_occupancy(raw, point_mask):
    """Renderer.py:189-190 -- occupancy := -100 where the sample has no neighbours, done
    in-place under no_grad in the reference: the VALUE changes, the gradient path does not."""
    occ = raw[:, 3]
    forced = torch.full_like(occ, -100.0)
    new_occ = torch.where(point_mask, occ, forced + (occ - occ.detach()))   # value -100, d/d occ = 1
    return torch.cat([raw[:, :3], new_occ[:, None]], -1)


def render_batch_ray(P, rays_d, rays_o, gt_depth, stage, cloud_pos, geo_feats, col_feats,
                     S=5, is_tracker=False, radius_query=0.08, dynamic_r_query=None,
                     rand_geo=None, rand_col=None, coef=0.1, encode_rel_pos=True,
                     exposure_mode='none', exposure_feat=None, sample_near_pcl=False,
                     near_surface=0.98, far_surface=1.02, near_end=0.3, tree=None, knn=None,
                     z_zero_depth=None, mask_not_near=None, return_aux=False, weighting='distance'):
    """Renderer.render_batch_ray, src/utils/Renderer.py:77-202."""
    R = rays_o.shape[0]
    z_vals, nz = sample_z_vals(gt_depth.detach(), S, near_surface, far_surface, near_end, z_zero_depth)
    p = sample_points(rays_o, rays_d, z_vals)
    dyn = None if dynamic_r_query is None else dynamic_r_query.reshape(-1, 1).repeat_interleave(S, 0)
    if knn is None:
        knn = find_neighbors(cloud_pos, p.detach(), radius_query, dyn, tree=tree)
    r2 = radius_sq_f64(p.shape[0], radius_query, dyn)
    if rand_geo is None:
        rand_geo = torch.zeros(C_DIM)
    if rand_col is None:
        rand_col = torch.zeros(C_DIM)
    raw, ray_mask, point_mask = point_forward(P, p, stage, knn, r2, geo_feats, col_feats, cloud_pos, S,
                                              is_tracker, rand_geo, rand_col, encode_rel_pos,
                                              exposure_mode, exposure_feat, weighting)
    raw = mask_occupancy(raw, point_mask).reshape(R, S, 4)
    depth, var, color, w = composite(raw, z_vals, coef)
    near_mask = torch.ones(R, dtype=torch.bool)
    if mask_not_near is not None:                                    # :158-163
        idx = torch.nonzero(~nz, as_tuple=True)[0][mask_not_near]
        near_mask[idx] = False
    valid = ray_mask & near_mask                                     # :198
    if not sample_near_pcl:
        depth = torch.where(nz, depth, torch.zeros_like(depth))     # :200-201
    if return_aux:
        return depth, var, color, valid, dict(z_vals=z_vals, p=p, knn=knn, raw=raw, weights=w,
                                              point_mask=point_mask)
    return depth, var, color, valid


# --------------------------------------------------------------------------- #
# losses of the two callers (used only to make scalar objectives for gradient parity)
# --------------------------------------------------------------------------- #
def tracker_loss(depth, var, color, gt_depth, gt_color, w_color=0.5):
    """src/Tracker.py:158-180 (handle_dynamic=True branch)."""
    var = var.detach()
    ok = (~torch.isnan(depth)) & (~torch.isnan(var))
    tmp = torch.abs(gt_depth - depth) / torch.sqrt(var + 1e-10)
    mask = (tmp < 10 * tmp.mean()) & (gt_depth > 0) & ok
    geo = torch.clamp(torch.abs(gt_depth - depth) / torch.sqrt(var + 1e-10), min=0.0, max=1e3)[mask].sum()
    col = torch.abs(gt_color - color)[mask].sum()
    return geo + w_color * col


def mapper_loss(depth, color, valid, gt_depth, gt_color, stage, w_color=0.1):
    """src/Mapper.py:521-553 (no exposure)."""
    m = (gt_depth > 0) & valid & (~torch.isnan(depth))
    loss = torch.abs(gt_depth[m] - depth[m]).sum()
    if stage == 'color':
        loss = loss + w_color * torch.abs(gt_color[m] - color[m]).sum()
    return loss


# --------------------------------------------------------------------------- #
# zero-depth rays: sample near the point cloud                                #
# --------------------------------------------------------------------------- #
def sample_near_pcl(cloud_pos, rays_o, rays_d, near, far, num, radius_query=0.08, tree=None):
    """NeuralPointCloud.sample_near_pcl, src/neural_point.py:217-277."""
    n_rays = rays_d.shape[0]
    intervals = 25
    zs = torch.linspace(near, float(far), steps=intervals)
    pts = (rays_o[..., None, :] + rays_d[..., None, :] * zs[..., :, None]).reshape(-1, 3)
    _, _, nnum = find_neighbors(cloud_pos, pts, radius_query, None, tree=tree)
    hit = nnum.numpy().reshape(n_rays, -1).astype(bool)
    invalid = hit.sum(-1) < 2
    sec = np.linspace(near, float(far), intervals)
    z_total = np.tile(np.linspace(near, float(far), num), (n_rays, 1))
    for r in np.nonzero(~invalid)[0]:
        cols = np.nonzero(hit[r])[0]
        z_total[r] = np.linspace(sec[cols[0]], sec[cols[1]], num=num)
    return torch.from_numpy(z_total).float(), torch.from_numpy(invalid)


# --------------------------------------------------------------------------- #
# point insertion filter                                                      #
# --------------------------------------------------------------------------- #
def add_points(cloud_pos, rays_o, rays_d, gt_depth, radius_add=0.04, dynamic_radius=None, is_pts_grad=False,
               radius_min=0.02, N_add=3, near_surface=0.98, far_surface=1.02, tree=None):
    """NeuralPointCloud.add_neural_points geometry, src/neural_point.py:107-145.
    Returns (kept_mask over depth>0 rays, new points (3*kept,3))."""
    m = gt_depth > 0
    o, d, D = rays_o[m], rays_d[m], gt_depth[m]
    dyn = None if dynamic_radius is None else dynamic_radius[m]
    pts_gt = (o[..., None, :] + d[..., None, :] * D[..., None, None]).reshape(-1, 3)
    keep = torch.ones(pts_gt.shape[0], dtype=torch.bool)
    if cloud_pos is not None and cloud_pos.shape[0] > 0:
        r = radius_min if is_pts_grad else radius_add
        _, _, n = find_neighbors(cloud_pos, pts_gt, r, dyn, tree=tree)
        keep = n == 0
    Ds = D.unsqueeze(-1).repeat(1, N_add)
    t = torch.linspace(0.0, 1.0, steps=N_add)
    z = near_surface * Ds * (1. - t) + far_surface * Ds * t
    pts = (o[..., None, :] + d[..., None, :] * z[..., :, None])[keep].reshape(-1, 3)
    return keep, pts


# ---------------------------------------------------------------------------------------------------------------------
# frustum feature selection
# ---------------------------------------------------------------------------------------------------------------------
def remap_bilinear_f32(img: np.ndarray, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """cv2.remap(img, x, y, INTER_LINEAR) for a float32 image and float32 maps, border constant 0, restated from
    OpenCV's imgproc/src/imgwarp.cpp (RemapInvoker + remapBilinear<Cast<float,float>, ..., float>): coordinates are
    rounded to 1/32 pixel (cvRound = round-half-even of x*32, non-finite / overflowing values -> INT_MIN), the integer
    part saturates to int16, weights are the exact products (1-fy)(1-fx) ... of the 1/32 fractions, the four taps
    (0 outside the image) are accumulated left to right in float32.  Bit-identical to OpenCV 4.13 on the inputs of
    `oracle/make_golden.py:run_frustum` and on 200 000 random coordinates incl. NaN/inf (checked when this was written)."""
    H, W = img.shape
    f32 = np.float32
    with np.errstate(all='ignore'):
        sx, sy = np.rint(x.astype(f32) * f32(32)), np.rint(y.astype(f32) * f32(32))

    def to_int(a):
        bad = ~np.isfinite(a) | (a >= 2.0 ** 31) | (a < -2.0 ** 31)
        return np.where(bad, -2.0 ** 31, a).astype(np.int64)
    sx, sy = to_int(sx), to_int(sy)
    fx, fy = (sx & 31).astype(f32) / f32(32), (sy & 31).astype(f32) / f32(32)
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        return np.where(ok, img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0))
    one = f32(1)
    w0, w1, w2, w3 = (one - fy) * (one - fx), (one - fy) * fx, fy * (one - fx), fy * fx
    return ((tap(iy, ix) * w0 + tap(iy, ix + 1) * w1) + tap(iy + 1, ix) * w2) + tap(iy + 1, ix + 1) * w3


def frustum_project(cloud_pos: np.ndarray, w2c: np.ndarray, fx, fy, cx, cy):
    """Mapper.get_mask_from_c2w, src/Mapper.py:131-147: float64 projection of the cloud (positions come from a Python
    list, hence float64) with the float32 inverse pose promoted to float64.  -> (u f32, v f32, z f64)."""
    p = np.asarray(cloud_pos, np.float64).reshape(-1, 3)
    w = np.asarray(w2c, np.float64)
    cam = [((w[r, 0] * p[:, 0] + w[r, 1] * p[:, 1]) + w[r, 2] * p[:, 2]) + w[r, 3] for r in range(3)]
    xc = -cam[0]                                              # :142 flip of the x axis
    u = (fx * xc + 0.0 * cam[1]) + cx * cam[2]
    v = (0.0 * xc + fy * cam[1]) + cy * cam[2]
    z = cam[2] + 1e-5                                         # :144  (uv[:, -1:] = 1*z_cam)
    with np.errstate(all='ignore'):
        return (u / z).astype(np.float32), (v / z).astype(np.float32), z


def frustum_indices(cloud_pos, c2w_f32, depth, H, W, fx, fy, cx, cy, edge=-4):
    """Mapper.get_mask_from_c2w (src/Mapper.py:120-168): indices of the points inside the (edge-enlarged) image whose
    camera depth is in [0, sampled sensor depth + 0.5]; zero sensor depth counts as the frame's largest sampled depth."""
    w2c = np.linalg.inv(np.asarray(c2w_f32, np.float32))      # :133, float32 LAPACK inverse like the reference
    u, v, z = frustum_project(cloud_pos, w2c, fx, fy, cx, cy)
    d = remap_bilinear_f32(np.asarray(depth, np.float32), u, v)
    mask = (u < W - edge) & (u > edge) & (v < H - edge) & (v > edge)
    d = d.copy()
    if d.size:
        d[d == 0] = np.max(d)                                 # :158-159
    mask &= (0 <= -z) & (-z <= d + 0.5)                       # float64 vs float32 -> compared in float64
    return np.nonzero(mask)[0]
