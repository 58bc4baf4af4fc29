"""The CPU restatements of the reference's CALLERS (point_slam_b200/iteration.py on top of oracle/point_slam_oracle.py -- what
bench.py's reference arm times) against vectors frozen from the unmodified `Tracker.optimize_cam_in_batch`,
`Mapper.optimize_map` and `Renderer.render_img` (tests/golden/caller_*.npz, render_img_*.npz).  CPU only."""
import numpy as np
import torch

from oracle import point_slam_oracle as O
from point_slam_b200 import iteration as IT
from tests import callers as K
from tests import cases as C

INTR = C.INTR


def _params(requires_grad):
    P = C.load_params(False)
    return {k: v.clone().requires_grad_(requires_grad and k != 'color_decoder.embedder._B') for k, v in P.items()}


class _Decoders:
    """The two things the iteration shells touch on the decoder object: color_decoder.parameters() and the param dict."""

    def __init__(self, P):
        self.P = P
        outer = self

        class CD:
            def parameters(self_inner):
                return [v for k, v in outer.P.items() if k.startswith('color_decoder.') and v.requires_grad]
        self.color_decoder = CD()


def _render_fn(P, cloud, S, rand):
    def render(npc, decoders, rays_d, rays_o, device, stage, gt_depth=None, npc_geo_feats=None, npc_col_feats=None,
               is_tracker=False, cloud_pos=None, dynamic_r_query=None, exposure_feat=None):
        rg, rc = rand()
        return O.render_batch_ray(P, rays_d, rays_o, gt_depth, stage, cloud, npc_geo_feats, npc_col_feats, S=S, is_tracker=is_tracker,
                                  radius_query=0.08, dynamic_r_query=dynamic_r_query, rand_geo=rg, rand_col=rc, coef=0.1,
                                  encode_rel_pos=True)
    return render


def test_tracker_shell_matches_reference_optimize_cam_in_batch():
    g = K.load('caller_tracker')
    scene = C.load_scene()
    P = _params(False)
    depth = torch.from_numpy(K.full_image(g['depth_win']))
    color = torch.from_numpy(K.full_image(g['color_win']))
    rq = torch.from_numpy(K.full_radius(g['r_query_win']))
    cam0 = torch.from_numpy(g['cam0'])
    quad, T = cam0[:4].clone().requires_grad_(True), cam0[4:].clone().requires_grad_(True)
    lr = float(g['lr'])
    opt = torch.optim.Adam([{'params': [T], 'lr': lr}, {'params': [quad], 'lr': lr * 0.2}])
    edge = tuple(int(v) for v in g['edge'])
    it_holder = [0]
    render = _render_fn(P, scene['cloud'], 5, lambda: (torch.from_numpy(g[f'rand_geo{it_holder[0]}']), torch.from_numpy(g[f'rand_col{it_holder[0]}'])))
    for it in range(3):
        it_holder[0] = it
        cam = torch.cat([quad, T], 0)
        grads = {}
        orig_step = opt.step

        def step():
            grads['q'], grads['T'] = quad.grad.clone(), T.grad.clone()
            return orig_step()
        opt.step = step
        with K.replay_randint([g[f'pix{it}']]):
            loss, n = IT.tracker_iteration(render, None, None, cam, opt, color, depth, rq, INTR, int(g['n_pixels']), 'cpu',
                                           scene['geo_feats'], scene['col_feats'], scene['cloud'], edge=edge, w_color=float(g['w_color']))
        opt.step = orig_step
        assert n == int(g[f'n_rays{it}'])
        assert abs(float(loss) - float(g[f'loss{it}'])) / abs(float(g[f'loss{it}'])) < 2e-5
        assert C.rel_err(grads['q'], g[f'grad_quad{it}']) < 3e-4 and C.rel_err(grads['T'], g[f'grad_T{it}']) < 3e-4
        assert C.rel_err(torch.cat([quad, T]).detach(), g[f'cam_after{it}']) < 1e-6


def test_mapper_shell_matches_reference_optimize_map():
    g = K.load('caller_mapper')
    scene = C.load_scene()
    P = _params(True)
    dec = _Decoders(P)
    n_it = int(g['n_iters'])
    frames = [dict(color=torch.from_numpy(K.full_image(g['color_win'][k])), depth=torch.from_numpy(K.full_image(g['depth_win'][k])),
                   c2w=torch.from_numpy(g['c2w'][k]).float(), dyn_r_query=torch.from_numpy(K.full_radius(g['r_query_win'][k])))
              for k in range(3)]
    cur = frames[2]
    # the map update the reference made first: add_neural_points on the recorded pixels, then the frustum selection
    from point_slam_b200.src import common
    with K.replay_randint([g['pix_add']]):
        ro, rd, gd, gc, i, j = common.get_samples(0, INTR['H'], 0, INTR['W'], g['pix_add'].shape[0], INTR['fx'], INTR['fy'], INTR['cx'],
                                                   INTR['cy'], cur['c2w'], cur['depth'], cur['color'], 'cpu', depth_filter=True, return_index=True)
    r_add = torch.from_numpy(K.full_radius(g['r_add_win'], fill=0.08))[j, i]
    keep, new = O.add_points(scene['cloud'], ro, rd, gd, dynamic_radius=r_add)
    assert np.array_equal(new.numpy(), g['added_pos'])
    cloud = torch.cat([scene['cloud'], new], 0)
    geo_all = torch.cat([scene['geo_feats'], torch.from_numpy(g['added_geo'])], 0)
    col_all = torch.cat([scene['col_feats'], torch.from_numpy(g['added_col'])], 0)
    idx = O.frustum_indices(cloud.numpy(), g['c2w'][2].astype(np.float32), cur['depth'].numpy(), INTR['H'], INTR['W'], INTR['fx'],
                            INTR['fy'], INTR['cx'], INTR['cy'], edge=-4)
    assert np.array_equal(idx, g['indices'])
    idx = torch.from_numpy(idx)

    class Npc:
        def get_geo_feats(self):
            return geo_all

        def get_col_feats(self):
            return col_all
    state = IT.MapperState(Npc(), dec, idx)
    lrs = {'geometry': (0.001, 0.03, 0.0), 'color': (0.005, 0.005, 0.005)}            # configs/point_slam.yaml:76-84
    U = idx.shape[0]
    it_holder = [0]
    render = _render_fn(P, cloud, 5, lambda: (torch.from_numpy(g[f'rand_geo{it_holder[0]}']), torch.from_numpy(g[f'rand_col{it_holder[0]}'])))
    for it in range(n_it):
        it_holder[0] = it
        stage = str(g['stages'][it])
        for grp, lr in zip(state.optimizer.param_groups, lrs[stage]):
            grp['lr'] = lr
        grads = {}
        orig_step = state.optimizer.step

        def step():
            grads['geo'] = state.geo.grad.clone()
            grads['col'] = None if state.col.grad is None else state.col.grad.clone()
            grads['dec'] = {k: v.grad.clone() for k, v in P.items() if v.grad is not None and k.startswith('color_decoder.')}
            return orig_step()
        state.optimizer.step = step
        with K.replay_randint(list(g[f'pix{it}'])):
            loss, n = IT.mapper_iteration(render, None, dec, state, [frames[0], frames[1], cur], INTR, int(g['n_pixels']), 'cpu', stage,
                                          cloud, w_color=float(g['w_color']))
        state.optimizer.step = orig_step
        assert n == int(g[f'n_rays{it}'])
        assert abs(float(loss) - float(g[f'loss{it}'])) / abs(float(g[f'loss{it}'])) < 2e-5, (it, float(loss), float(g[f'loss{it}']))
        assert C.rel_err(grads['geo'], K.dense_rows(g[f'grad_geo_rows{it}'], g[f'grad_geo_vals{it}'], U)) < 2e-4
        if stage == 'color':
            assert C.rel_err(grads['col'], K.dense_rows(g[f'grad_col_rows{it}'], g[f'grad_col_vals{it}'], U)) < 2e-4
            n_dec = 0
            for k, v in g.items():
                if k.startswith(f'grad_dec{it}.'):
                    assert C.rel_err(grads['dec']['color_decoder.' + k.split('.', 1)[1]], v) < 3e-4, k
                    n_dec += 1
            assert n_dec >= 20
        else:
            assert grads['col'] is None and not grads['dec']
    # end state: rows that moved and the colour decoder
    # (Adam's m / (sqrt(v) + eps) amplifies rounding differences of tiny gradients: 1e-4 of the feature scale)
    want = geo_all[idx].clone(); want[torch.from_numpy(g['geo_after_rows'])] = torch.from_numpy(g['geo_after_vals'])
    assert C.rel_err(state.npc_geo[idx], want) < 1e-4
    want = col_all[idx].clone(); want[torch.from_numpy(g['col_after_rows'])] = torch.from_numpy(g['col_after_vals'])
    assert C.rel_err(state.npc_col[idx], want) < 1e-4
    n_dec = 0
    for k, v in g.items():
        if k.startswith('dec_after.'):
            assert C.rel_err(P['color_decoder.' + k.split('.', 1)[1]].detach(), v) < 1e-4, k
            n_dec += 1
    assert n_dec >= 20


def _render_img_oracle(g, sample_near_pcl, encode_rel_pos):
    scene = C.load_scene()
    P = C.load_params(False)
    H, W, fx, fy, cx, cy = g['intr']
    H, W = int(H), int(W)
    c2w = torch.from_numpy(g['c2w']).float()
    ro, rd = O.rays_full_image(H, W, fx, fy, cx, cy, c2w)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    gd = torch.from_numpy(g['gt_depth']).reshape(-1)
    rq = torch.from_numpy(g['r_query']).reshape(-1)
    outs = []
    from scipy.spatial import cKDTree
    tree = cKDTree(scene['cloud'].double().numpy())
    for s in range(0, H * W, 3000):                      # ray_batch_size, Renderer.py:7,248
        sl = slice(s, s + 3000)
        d_b = gd[sl]
        zz = mnn = None
        if sample_near_pcl and bool((d_b <= 0).any()):
            far = torch.minimum(5 * d_b.mean(), torch.max(d_b * 1.2)).float()
            zero = d_b <= 0
            zz, mnn = O.sample_near_pcl(scene['cloud'], ro[sl][zero], rd[sl][zero], 0.3, far, 5, tree=tree)
        outs.append(O.render_batch_ray(P, rd[sl], ro[sl], d_b, 'color', scene['cloud'], scene['geo_feats'], scene['col_feats'], S=5,
                                       dynamic_r_query=rq[sl], rand_geo=torch.from_numpy(g['rand_geo']), rand_col=torch.from_numpy(g['rand_col']),
                                       encode_rel_pos=encode_rel_pos, sample_near_pcl=sample_near_pcl, z_zero_depth=zz, mask_not_near=mnn,
                                       tree=tree)[:3])
    depth = torch.cat([o[0] for o in outs]).reshape(H, W)
    unc = torch.cat([o[1] for o in outs]).reshape(H, W)
    color = torch.cat([o[2] for o in outs]).reshape(H, W, 3)
    return depth, unc, color


def test_render_img_oracle_matches_reference():
    for name, near_pcl, rel in (('replica', False, True), ('tum', True, False)):
        g = K.load(f'render_img_{name}')
        with torch.no_grad():
            d, u, c = _render_img_oracle(g, near_pcl, rel)
        assert C.rel_err(d, g['depth']) < 2e-5, name
        assert C.rel_err(c, g['color']) < 2e-5, name
        assert C.rel_err(u, g['uncertainty']) < 5e-4, name
        if not near_pcl:
            assert bool((d.reshape(-1)[torch.from_numpy(g['gt_depth']).reshape(-1) <= 0] == 0).all())      # Renderer.py:200-201
