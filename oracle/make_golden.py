"""Freeze outputs of the UNMODIFIED reference as golden vectors  --  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py            # writes tests/golden/*.npz  (build container only)

Every array in the written files was produced by code imported from
/root/reference through `oracle/ref_harness.py` (exact-kNN faiss stand-in, CPU,
float32).  The files are the pin for `oracle/point_slam_oracle.py`
(tests/test_oracle_vs_golden.py) and, through it, for the CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H            # noqa: E402
from point_slam_b200 import synth              # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
INTR = synth.TUM_INTRINSICS


def fp16_exact(x):
    return x.astype(np.float16).astype(np.float32)


def make_scene():
    """Camera 1.3 m from a wall + table corner; cloud cropped to the viewed window; one hole."""
    c2w = synth.look_at([1.6, 2.6, 1.1], [1.5, 0.0, 0.7])
    depth, color = synth.make_frame(c2w, INTR)
    _, r_query = synth.sobel_radius_map(color)
    rng = np.random.default_rng(1219)
    # pixel window (keeps the cropped cloud small)
    j0, j1, i0, i1 = 150, 330, 220, 420
    o, d = synth.pixel_rays(c2w, INTR['H'], INTR['W'], INTR['fx'], INTR['fy'], INTR['cx'], INTR['cy'])
    win = (o + d[j0:j1, i0:i1] * depth[j0:j1, i0:i1, None]).reshape(-1, 3)
    lo, hi = win.min(0) - 0.30, win.max(0) + 0.30
    cloud = synth.make_cloud(500000)
    keep = np.all((cloud >= lo) & (cloud <= hi), 1)
    cloud = cloud[keep]
    # punch a hole so that some samples have < 2 neighbours
    hole_c = (o + d[200, 300] * depth[200, 300]).astype(np.float32)
    cloud = cloud[np.linalg.norm(cloud - hole_c, axis=1) > 0.22]
    gf, cf = synth.make_features(cloud.shape[0])
    return dict(c2w=c2w, depth=depth, color=color, r_query=r_query, window=(j0, j1, i0, i1),
                cloud=cloud, geo_feats=fp16_exact(gf), col_feats=fp16_exact(cf), hole=hole_c, rng=rng)


def pick_pixels(scene, n, seed, include_hole=True):
    j0, j1, i0, i1 = scene['window']
    rng = np.random.default_rng(seed)
    jj = rng.integers(j0, j1, n)
    ii = rng.integers(i0, i1, n)
    if include_hole:                       # a cluster of pixels looking into the hole
        k = n // 8
        jj[:k] = 200 + rng.integers(-6, 7, k)
        ii[:k] = 300 + rng.integers(-6, 7, k)
    return jj, ii


def sparse_rows(g):
    g = g.detach().numpy()
    rows = np.nonzero(np.abs(g).sum(1))[0]
    return rows.astype(np.int64), g[rows]


def grads_dict(model, prefix='grad.'):
    out = {}
    for k, p in model.named_parameters():
        if p.grad is not None:
            out[prefix + k] = p.grad.detach().numpy().copy()
    return out


def run_case(name, scene, config, overrides, stage, is_tracker, R, S=None, seed=7, zero_depth_frac=0.0,
             loss_kind='mapper', use_cam_tensor=False, exposure=None, pretrained=True, write_params=True):
    ov = dict(overrides or {})
    if S is not None:
        ov['rendering.N_surface'] = S
    ref = H.load_reference(config, ov)
    cfg = ref['cfg']
    S = cfg['rendering']['N_surface']
    model = H.build_decoders(ref, pretrained=pretrained)
    P = H.decoder_params(model)
    npc = H.build_npc(ref, scene['cloud'], scene['geo_feats'], scene['col_feats'])
    renderer = H.build_renderer(ref, INTR)
    common = ref['common']

    jj, ii = pick_pixels(scene, R, seed)
    i_t = torch.from_numpy(ii).float()
    j_t = torch.from_numpy(jj).float()
    gt_depth = torch.from_numpy(scene['depth'][jj, ii].copy())
    gt_color = torch.from_numpy(scene['color'][jj, ii].copy())
    if zero_depth_frac > 0:
        nz = int(R * zero_depth_frac)
        gt_depth[-nz:] = 0.0
    dyn = torch.from_numpy(scene['r_query'][jj, ii].copy()) if cfg['use_dynamic_radius'] else None

    c2w64 = scene['c2w']
    save = dict(pix_i=ii, pix_j=jj, gt_depth=gt_depth.numpy(), gt_color=gt_color.numpy(), S=S,
                stage=stage, is_tracker=is_tracker, loss_kind=loss_kind,
                use_dynamic_radius=bool(cfg['use_dynamic_radius']),
                radius_query=cfg['pointcloud']['radius_query'],
                encode_rel_pos=bool(cfg['model']['encode_rel_pos_in_col']),
                encode_exposure=bool(cfg['model']['encode_exposure']),
                sample_near_pcl=bool(cfg['rendering']['sample_near_pcl']),
                near_end=cfg['rendering']['near_end'], seed=seed, nn_weighting=cfg['pointcloud']['nn_weighting'])
    if dyn is not None:
        save['dynamic_r_query'] = dyn.numpy()

    cam = None
    if use_cam_tensor:
        from scipy.spatial.transform import Rotation
        q = np.roll(Rotation.from_matrix(c2w64[:3, :3]).as_quat(), 1)
        cam = torch.tensor(np.concatenate([q, c2w64[:3, 3]]), dtype=torch.float32, requires_grad=True)
        c2w = common.get_camera_from_tensor(cam)
        save['cam_tensor'] = cam.detach().numpy()
    else:
        c2w = torch.tensor(c2w64[:3, :4], dtype=torch.float32)
        save['c2w'] = c2w.numpy()
    rays_o, rays_d = common.get_rays_from_uv(i_t, j_t, c2w, INTR['fx'], INTR['fy'], INTR['cx'], INTR['cy'], 'cpu')
    save['rays_o'] = rays_o.detach().numpy().copy()
    save['rays_d'] = rays_d.detach().numpy().copy()

    geo = npc.get_geo_feats().clone().requires_grad_(True)
    col = npc.get_col_feats().clone().requires_grad_(True)
    cloud_t = torch.tensor(npc.cloud_pos()).reshape(-1, 3)
    exposure_feat = None
    if exposure == 'feat':
        torch.manual_seed(99)
        exposure_feat = (torch.randn(cfg['model']['exposure_dim']) * 0.5).requires_grad_(True)
        save['exposure_feat'] = exposure_feat.detach().numpy()

    ra, rb = H.draw_rand_vecs(seed)
    save['rand_geo'], save['rand_col'] = ra.numpy(), rb.numpy()
    torch.manual_seed(seed)
    depth, var, color, valid = renderer.render_batch_ray(
        npc, model, rays_d, rays_o, 'cpu', stage, gt_depth=gt_depth, npc_geo_feats=geo, npc_col_feats=col,
        is_tracker=is_tracker, cloud_pos=cloud_t, dynamic_r_query=dyn, exposure_feat=exposure_feat)
    save.update(depth=depth.detach().numpy(), var=var.detach().numpy(), color=color.detach().numpy(),
                valid=valid.numpy())

    # recover z_vals / kNN / raw through the same reference entry points (no second implementation)
    with torch.no_grad():
        far = torch.minimum(5 * gt_depth.mean(), torch.max(gt_depth * 1.2))
        nzm = gt_depth > 0
        if (~nzm).any() and cfg['rendering']['sample_near_pcl']:
            zz, inv = npc.sample_near_pcl(rays_o[~nzm].detach().clone(), rays_d[~nzm].detach().clone(),
                                          cfg['rendering']['near_end'], torch.max(far.repeat(R, 1).float()), S)
            save['z_zero_depth'], save['mask_not_near'] = zz.numpy(), inv.numpy()

    if loss_kind == 'tracker':
        unc = var.detach()
        nan_mask = (~torch.isnan(depth)) & (~torch.isnan(unc))
        tmp = torch.abs(gt_depth - depth) / torch.sqrt(unc + 1e-10)
        mask = (tmp < 10 * tmp.mean()) & (gt_depth > 0) & nan_mask
        loss = torch.clamp(torch.abs(gt_depth - depth) / torch.sqrt(unc + 1e-10), min=0.0, max=1e3)[mask].sum()
        loss = loss + 0.5 * torch.abs(gt_color - color)[mask].sum()
    else:
        m = (gt_depth > 0) & valid & (~torch.isnan(depth))
        loss = torch.abs(gt_depth[m] - depth[m]).sum()
        if stage == 'color':
            loss = loss + 0.1 * torch.abs(gt_color[m] - color[m]).sum()
    loss.backward()
    save['loss'] = np.float32(loss.item())
    if cam is not None:
        save['grad.cam_tensor'] = cam.grad.numpy().copy()
    if exposure_feat is not None and exposure_feat.grad is not None:
        save['grad.exposure_feat'] = exposure_feat.grad.numpy().copy()
    if geo.grad is not None:
        save['grad.geo_rows'], save['grad.geo_vals'] = sparse_rows(geo.grad)
    if col.grad is not None:
        save['grad.col_rows'], save['grad.col_vals'] = sparse_rows(col.grad)
    save.update(grads_dict(model))
    np.savez_compressed(os.path.join(OUT, f'case_{name}.npz'), **save)
    pkey = f"decoders_{'exposure' if cfg['model']['encode_exposure'] else 'base'}.npz"
    if write_params:
        np.savez_compressed(os.path.join(OUT, pkey), **{k: v.numpy() for k, v in P.items()})
    print(f'{name:28s} R={R} S={S} loss={loss.item():.6f} valid={int(valid.sum())}/{R} '
          f'depth[{depth.min().item():.3f},{depth.max().item():.3f}]')


def run_aux(scene):
    """Small stand-alone vectors: kNN, get_rays, composite, add_neural_points, sample_near_pcl."""
    ref = H.load_reference('configs/Replica/room0.yaml')
    common = ref['common']
    out = {}
    # get_rays on a tiny camera
    c2w = torch.tensor(scene['c2w'][:3, :4], dtype=torch.float32)
    ro, rd = common.get_rays(12, 16, 20.0, 21.0, 7.5, 5.5, c2w, 'cpu')
    out['get_rays_o'], out['get_rays_d'] = ro.numpy().copy(), rd.numpy().copy()
    # composite on random raw
    g = torch.Generator().manual_seed(5)
    raw = torch.randn(64, 5, 4, generator=g) * 8
    raw[::7, :, 3] = -100.0
    z = torch.sort(torch.rand(64, 5, generator=g) * 3 + 0.5, -1)[0]
    out['comp_raw'], out['comp_z'] = raw.numpy().copy(), z.numpy().copy()
    d, v, c, w = common.raw2outputs_nerf_color(raw.clone(), z, torch.randn(64, 3), device='cpu', coef=0.1)
    out['comp_depth'], out['comp_var'], out['comp_rgb'], out['comp_w'] = d.numpy(), v.numpy(), c.numpy(), w.numpy()
    # add_neural_points on top of the scene cloud (dynamic radius, then gradient-picked fixed radius_min)
    npc = H.build_npc(ref, scene['cloud'], scene['geo_feats'], scene['col_feats'])
    jj, ii = pick_pixels(scene, 400, 21)
    i_t, j_t = torch.from_numpy(ii).float(), torch.from_numpy(jj).float()
    rays_o, rays_d = common.get_rays_from_uv(i_t, j_t, c2w, INTR['fx'], INTR['fy'], INTR['cx'], INTR['cy'], 'cpu')
    gd = torch.from_numpy(scene['depth'][jj, ii].copy())
    gd[::19] = 0.0
    gc = torch.from_numpy(scene['color'][jj, ii].copy())
    r_add = torch.from_numpy((scene['r_query'][jj, ii] / 2.0).copy())
    n0 = npc.pts_num()
    torch.manual_seed(3)
    k1 = npc.add_neural_points(rays_o, rays_d, gd, gc, dynamic_radius=r_add[gd > 0])
    n1 = npc.pts_num()
    k2 = npc.add_neural_points(rays_o, rays_d, gd, gc, is_pts_grad=True)
    n2 = npc.pts_num()
    cp = np.asarray(npc.cloud_pos(), dtype=np.float32)
    out.update(add_pix_i=ii, add_pix_j=jj, add_rays_o=rays_o.numpy().copy(), add_rays_d=rays_d.numpy().copy(),
               add_depth=gd.numpy(), add_r_add=r_add.numpy(), add_kept1=np.int64(int(k1)), add_kept2=np.int64(int(k2)),
               add_new1=cp[n0:n1], add_new2=cp[n1:n2], add_input_pos=np.asarray(npc.input_pos(), np.float32))
    # sample_near_pcl on rays of the window (reference NeuralPointCloud method, radius_query fixed 0.08)
    npc2 = H.build_npc(ref, scene['cloud'], scene['geo_feats'], scene['col_feats'])
    zz, inv = npc2.sample_near_pcl(rays_o[:96].clone(), rays_d[:96].clone(), 0.3, torch.tensor(4.2), 5)
    out['snp_z'], out['snp_invalid'] = zz.numpy(), inv.numpy()
    # kNN contract on sample points (find_neighbors_faiss of the reference class over the exact index)
    pts = (rays_o[:64, None, :] + rays_d[:64, None, :] * (gd[:64, None, None].clamp(min=0.5) *
                                                         torch.linspace(0.98, 1.02, 5)[None, :, None])).reshape(-1, 3)
    D, I, n = npc2.find_neighbors_faiss(pts, step='query')
    out['knn_pts'], out['knn_D'], out['knn_I'], out['knn_n'] = pts.numpy(), D.numpy(), I.numpy(), n.numpy()
    np.savez_compressed(os.path.join(OUT, 'aux.npz'), **out)
    print('aux: add kept', int(k1), int(k2), 'near_pcl invalid', int(inv.sum()), '/ 96')


def run_expo(scene):
    rep = 'configs/Replica/room0.yaml'
    ov = {'pointcloud.nn_weighting': 'expo'}
    # mapper only: with a pose gradient the reference itself raises (decoder.py:156-157 zeroes the output of torch.exp in
    # place, which ExpBackward needs) -- 'expo' cannot be trained through the tracker path in the reference
    run_case('expo_mapper', scene, rep, ov, 'color', False, 96, seed=17, write_params=False)
    run_case('expo_mapper_geometry', scene, rep, ov, 'geometry', False, 96, seed=18, write_params=False)


def run_frustum():
    """Mapper.get_mask_from_c2w (src/Mapper.py:120-168) of the unmodified reference on a whole-room cloud: two poses of a
    quarter-size camera, sensor depth with zero-depth holes, the shipped frustum_edge (-4) and a positive one."""
    import types
    Mapper = H.load_mapper_class()
    intr = dict(H=120, W=160, fx=INTR['fx'] / 4, fy=INTR['fy'] / 4, cx=INTR['cx'] / 4, cy=INTR['cy'] / 4)
    cloud = synth.make_cloud(30000, seed=77)
    rng = np.random.default_rng(78)
    extra = rng.uniform([-1, -1, -1], [7, 5, 3.7], (2000, 3)).astype(np.float32)      # outside the walls / behind surfaces
    cloud = np.concatenate([cloud, extra], 0)
    out = dict(cloud=cloud, intr=np.array([intr[k] for k in ('H', 'W', 'fx', 'fy', 'cx', 'cy')], np.float64))
    poses = [synth.look_at([1.6, 2.6, 1.1], [1.5, 0.0, 0.7]), synth.trajectory(7, seed=5)[3]]
    for k, (c2w, edge) in enumerate(zip(poses, (-4, 10))):
        depth, _ = synth.make_frame(c2w, intr)
        depth = depth.astype(np.float32)
        depth[rng.random(depth.shape) < 0.04] = 0.0
        depth[:9, :13] = 0.0                                   # a zero-depth corner: border taps mix with it
        npc = types.SimpleNamespace(cloud_pos=lambda c=cloud: c.tolist())
        me = types.SimpleNamespace(H=intr['H'], W=intr['W'], fx=intr['fx'], fy=intr['fy'], cx=intr['cx'], cy=intr['cy'],
                                   npc=npc, frustum_edge=edge)
        c2w_t = torch.from_numpy(c2w).float()
        idx = Mapper.get_mask_from_c2w(me, c2w_t, depth)
        out[f'c2w{k}'], out[f'depth{k}'], out[f'edge{k}'] = c2w_t.numpy(), depth, np.int64(edge)
        out[f'indices{k}'] = np.asarray(idx, np.int64)
        print(f'frustum {k}: {len(idx)} of {cloud.shape[0]} points selected (edge {edge})')
    np.savez_compressed(os.path.join(OUT, 'frustum.npz'), **out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'frustum':        # only that fixture (the others stay byte-identical)
        run_frustum()
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'expo':           # nn_weighting='expo' (decoder.py:154-156; unused by the shipped configs)
        run_expo(make_scene())
        return
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    scene = make_scene()
    print('scene cloud', scene['cloud'].shape)
    np.savez_compressed(os.path.join(OUT, 'scene.npz'), cloud=scene['cloud'],
                        geo_feats=scene['geo_feats'].astype(np.float16), col_feats=scene['col_feats'].astype(np.float16),
                        c2w=scene['c2w'], hole=scene['hole'])
    rep, tum, scn = 'configs/Replica/room0.yaml', 'configs/TUM_RGBD/freiburg1_desk.yaml', 'configs/ScanNet/scene0000.yaml'
    run_case('mapper_color', scene, rep, None, 'color', False, 160)
    run_case('mapper_geometry', scene, rep, None, 'geometry', False, 160, seed=8)
    run_case('tracker_color', scene, rep, None, 'color', True, 160, seed=9, loss_kind='tracker', use_cam_tensor=True)
    run_case('fixed_radius_zero_depth', scene, rep, {'use_dynamic_radius': False}, 'color', False, 128, seed=10,
             zero_depth_frac=0.1)
    run_case('tum_near_pcl', scene, tum, None, 'color', False, 128, seed=11, zero_depth_frac=0.15)
    run_case('tum_tracker', scene, tum, None, 'color', True, 128, seed=12, loss_kind='tracker', use_cam_tensor=True)
    run_case('s32_color', scene, rep, None, 'color', False, 24, S=32, seed=13)
    run_case('exposure_tracker', scene, scn, None, 'color', True, 96, seed=14, loss_kind='tracker',
             use_cam_tensor=True, exposure='feat')
    run_case('exposure_mapper_raw', scene, scn, None, 'color', False, 96, seed=15)
    run_aux(scene)
    run_expo(scene)
    run_frustum()


if __name__ == '__main__':
    main()
