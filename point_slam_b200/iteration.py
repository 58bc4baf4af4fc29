"""The two callers' inner iterations, restated so that bench.py / tests can drive the hot path the way the reference
does: `tracker_iteration` == Tracker.optimize_cam_in_batch (src/Tracker.py:89-186, Replica branch: uniform pixel
sampling, handle_dynamic loss, Adam on [quat, T]); `mapper_iteration` == one pass of the joint loop of
Mapper.optimize_map (src/Mapper.py:408-568: per-keyframe pixel sampling, frustum-selected feature slices written
into the full tensors with index_put, stage-dependent loss, Adam step, write-back).

Both take a `render` callable with Renderer.render_batch_ray's signature so the same shell times the CUDA path
(point_slam_b200.src.utils.Renderer) and, in bench.py's reference arm only, the CPU oracle.
"""
import torch

from .src import common


def tracker_iteration(render, npc, decoders, cam, optimizer, gt_color, gt_depth, dyn_r_query, intr, n_pixels, device,
                      geo_feats, col_feats, cloud_pos, edge=(20, 20), w_color=0.5, exposure_feat=None):
    H, W = intr['H'], intr['W']
    optimizer.zero_grad()
    c2w = common.get_camera_from_tensor(cam)
    rays_o, rays_d, b_depth, b_color, i, j = common.get_samples(
        edge[0], H - edge[0], edge[1], W - edge[1], n_pixels, intr['fx'], intr['fy'], intr['cx'], intr['cy'], c2w,
        gt_depth, gt_color, device, depth_filter=True, return_index=True)
    b_rq = dyn_r_query[j, i] if dyn_r_query is not None else None
    with torch.no_grad():
        inside = b_depth <= torch.minimum(10 * b_depth.median(), 1.2 * torch.max(b_depth))
    rays_d, rays_o, b_depth, b_color = rays_d[inside], rays_o[inside], b_depth[inside], b_color[inside]
    b_rq = b_rq[inside] if b_rq is not None else None
    depth, unc, color, _ = render(npc, decoders, rays_d, rays_o, device, stage='color', gt_depth=b_depth,
                                  npc_geo_feats=geo_feats, npc_col_feats=col_feats, is_tracker=True,
                                  cloud_pos=cloud_pos, dynamic_r_query=b_rq, exposure_feat=exposure_feat)
    unc = unc.detach()
    ok = (~torch.isnan(depth)) & (~torch.isnan(unc))
    tmp = torch.abs(b_depth - depth) / torch.sqrt(unc + 1e-10)
    mask = (tmp < 10 * tmp.mean()) & (b_depth > 0) & ok
    loss = torch.clamp(tmp, min=0.0, max=1e3)[mask].sum() + w_color * torch.abs(b_color - color)[mask].sum()
    loss.backward()
    optimizer.step()
    optimizer.zero_grad()
    return loss.detach(), int(rays_o.shape[0])


def frustum_indices(cloud_pos, c2w, intr, margin=4):
    """Indices of the points that project inside the (slightly enlarged) image of pose c2w -- the selection
    Mapper.get_mask_from_c2w makes (src/Mapper.py:120-168), without the depth-consistency test."""
    R, t = c2w[:3, :3], c2w[:3, 3]
    pc = (cloud_pos - t) @ R                               # camera frame (x right, y up, z backward)
    z = -pc[:, 2]
    u = intr['fx'] * pc[:, 0] / z.clamp_min(1e-6) + intr['cx']
    v = -intr['fy'] * pc[:, 1] / z.clamp_min(1e-6) + intr['cy']
    m = (z > 0) & (u > -margin) & (u < intr['W'] + margin) & (v > -margin) & (v < intr['H'] + margin)
    return torch.nonzero(m, as_tuple=True)[0]


class MapperState:
    """Optimisable slices + Adam, set up like Mapper.optimize_map (src/Mapper.py:345-402)."""

    def __init__(self, npc, decoders, indices, lr_dec=0.005, lr_geo=0.005, lr_col=0.005, capturable=False):
        self.indices = indices
        self.npc_geo = npc.get_geo_feats().detach().clone()
        self.npc_col = npc.get_col_feats().detach().clone()
        self.geo = self.npc_geo[indices].detach().clone().requires_grad_(True)
        self.col = self.npc_col[indices].detach().clone().requires_grad_(True)
        self.optimizer = torch.optim.Adam([{'params': list(decoders.color_decoder.parameters()), 'lr': lr_dec},
                                           {'params': [self.geo], 'lr': lr_geo}, {'params': [self.col], 'lr': lr_col}],
                                          capturable=capturable)


def mapper_iteration(render, npc, decoders, state, keyframes, intr, n_pixels, device, stage, cloud_pos, w_color=0.1):
    """keyframes: list of dicts(color (H,W,3), depth (H,W), c2w (3x4 / 4x4), dyn_r_query (H,W) f64 or None)."""
    H, W = intr['H'], intr['W']
    idx = state.indices
    npc_geo, npc_col = state.npc_geo, state.npc_col
    npc_geo[idx] = state.geo                                # index_put that keeps the graph (Mapper.py:411-414)
    npc_col[idx] = state.col
    state.optimizer.zero_grad()
    per = n_pixels // len(keyframes)
    ro, rd, dep, colr, rq = [], [], [], [], []
    for kf in keyframes:
        o, d, gd, gc, i, j = common.get_samples(0, H, 0, W, per, intr['fx'], intr['fy'], intr['cx'], intr['cy'],
                                                kf['c2w'], kf['depth'], kf['color'], device, depth_filter=True,
                                                return_index=True)
        ro.append(o.float()); rd.append(d.float()); dep.append(gd.float()); colr.append(gc.float())
        if kf.get('dyn_r_query') is not None:
            rq.append(kf['dyn_r_query'][j, i])
    rays_o, rays_d, b_depth, b_color = torch.cat(ro), torch.cat(rd), torch.cat(dep), torch.cat(colr)
    b_rq = torch.cat(rq) if rq else None
    with torch.no_grad():
        inside = b_depth <= torch.minimum(10 * b_depth.median(), 1.2 * torch.max(b_depth))
    rays_d, rays_o, b_depth, b_color = rays_d[inside], rays_o[inside], b_depth[inside], b_color[inside]
    b_rq = b_rq[inside] if b_rq is not None else None
    depth, unc, color, valid = render(npc, decoders, rays_d, rays_o, device, stage, gt_depth=b_depth,
                                      npc_geo_feats=npc_geo, npc_col_feats=npc_col, is_tracker=False,
                                      cloud_pos=cloud_pos, dynamic_r_query=b_rq, exposure_feat=None)
    m = (b_depth > 0) & valid & (~torch.isnan(depth))
    loss = torch.abs(b_depth[m] - depth[m]).sum()
    if stage == 'color':
        loss = loss + w_color * torch.abs(b_color[m] - color[m]).sum()
    loss.backward()
    state.optimizer.step()
    state.optimizer.zero_grad()
    state.npc_geo, state.npc_col = npc_geo.detach(), npc_col.detach()          # Mapper.py:560-565
    state.npc_geo[idx], state.npc_col[idx] = state.geo.detach().clone(), state.col.detach().clone()
    return loss.detach(), int(rays_o.shape[0])
