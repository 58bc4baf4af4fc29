"""Run a golden case through the drop-in API (NeuralPointCloud + POINT + Renderer -> CUDA library) on cuda:0."""
import types

import numpy as np
import torch

from point_slam_b200.default_config import make_cfg
from point_slam_b200.src import common
from point_slam_b200.src.conv_onet import config as model_config
from point_slam_b200.src.neural_point import NeuralPointCloud
from point_slam_b200.src.utils.Renderer import Renderer
from tests import cases as C

DEV = 'cuda:0'


def case_cfg(c):
    ds = 'scannet' if c['encode_exposure'] else ('replica' if c['encode_rel_pos'] else 'tum')
    return make_cfg(ds, DEV, **{'use_dynamic_radius': c['use_dynamic_radius'], 'rendering.N_surface': c['S'],
                                'rendering.sample_near_pcl': c['sample_near_pcl'],
                                'pointcloud.radius_query': float(c['radius_query']),
                                'pointcloud.nn_weighting': c.get('nn_weighting', 'distance')})


def build_objects(c, scene=None):
    cfg = case_cfg(c)
    decoders = model_config.get_model(cfg).to(DEV)
    P = C.load_params(c['encode_exposure'])
    sd = {k: v for k, v in P.items() if k != 'color_decoder.embedder._B'}
    missing = decoders.load_state_dict(sd, strict=True)
    decoders.color_decoder.embedder._B = P['color_decoder.embedder._B'].to(DEV)
    scene = scene or C.load_scene()
    npc = NeuralPointCloud(cfg)
    npc._cloud_pos = scene['cloud']
    npc._pts_num = scene['cloud'].shape[0]
    npc.geo_feats = scene['geo_feats'].to(DEV).clone()
    npc.col_feats = scene['col_feats'].to(DEV).clone()
    npc.index.add(npc._pos)
    intr = C.INTR
    renderer = Renderer(cfg, None, types.SimpleNamespace(**{k: intr[k] for k in ('H', 'W', 'fx', 'fy', 'cx', 'cy')}))
    renderer.sigmoid_coefficient = 0.1
    return cfg, decoders, npc, renderer


def run_case_gpu(c, objects=None, freeze_decoders=False):
    cfg, decoders, npc, renderer = objects or build_objects(c)
    decoders.requires_grad_(not freeze_decoders)
    intr = C.INTR
    t = lambda k, dt=torch.float32: torch.from_numpy(np.asarray(c[k])).to(device=DEV, dtype=dt)
    gt_depth, gt_color = t('gt_depth'), t('gt_color')
    cam = None
    if 'cam_tensor' in c:
        cam = t('cam_tensor').requires_grad_(True)
        c2w = common.get_camera_from_tensor(cam)
        rays_o, rays_d = common.get_rays_from_uv(t('pix_i'), t('pix_j'), c2w, intr['fx'], intr['fy'], intr['cx'], intr['cy'], DEV)
    else:
        rays_o, rays_d = t('rays_o'), t('rays_d')
    dyn = t('dynamic_r_query', torch.float64) if 'dynamic_r_query' in c else None
    ef = t('exposure_feat').requires_grad_(True) if 'exposure_feat' in c else None
    geo = npc.get_geo_feats().clone().requires_grad_(True)
    col = npc.get_col_feats().clone().requires_grad_(True)
    rg, rc = t('rand_geo'), t('rand_col')
    decoders.draw_no_neighbor_vectors = lambda stage, device: (rg, rc if stage == 'color' else None)
    decoders.zero_grad()
    depth, var, color, valid = renderer.render_batch_ray(
        npc, decoders, rays_d, rays_o, DEV, c['stage'], gt_depth=gt_depth, npc_geo_feats=geo, npc_col_feats=col,
        is_tracker=c['is_tracker'], cloud_pos=npc.cloud_pos_tensor(), dynamic_r_query=dyn, exposure_feat=ef)
    if c['loss_kind'] == 'tracker':
        unc = var.detach()
        ok = (~torch.isnan(depth)) & (~torch.isnan(unc))
        tmp = torch.abs(gt_depth - depth) / torch.sqrt(unc + 1e-10)
        mask = (tmp < 10 * tmp.mean()) & (gt_depth > 0) & ok
        loss = torch.clamp(torch.abs(gt_depth - depth) / torch.sqrt(unc + 1e-10), min=0.0, max=1e3)[mask].sum()
        loss = loss + 0.5 * torch.abs(gt_color - color)[mask].sum()
    else:
        m = (gt_depth > 0) & valid & (~torch.isnan(depth))
        loss = torch.abs(gt_depth[m] - depth[m]).sum()
        if c['stage'] == 'color':
            loss = loss + 0.1 * torch.abs(gt_color[m] - color[m]).sum()
    loss.backward()
    out = dict(depth=depth.detach(), var=var.detach(), color=color.detach(), valid=valid, loss=loss.detach(),
               grad_geo=geo.grad, grad_col=col.grad,
               grad_params={k: p.grad for k, p in decoders.named_parameters() if p.grad is not None})
    if cam is not None:
        out['grad_cam'] = cam.grad
    if ef is not None:
        out['grad_exposure_feat'] = ef.grad
    return out
