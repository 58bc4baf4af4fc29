"""Freeze the reference's own CALLERS of the render path as golden vectors  --  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_callers.py        # writes tests/golden/caller_*.npz, render_img_*.npz  (build container only)

What runs is the UNMODIFIED reference, imported from /root/reference through `oracle/ref_harness.py`:
  * `Tracker.optimize_cam_in_batch` (src/Tracker.py:89-186), called unbound on a namespace that carries exactly the attributes
    the method reads, three iterations with the reference's separate_LR Adam (Tracker.py:289-299);
  * `Mapper.optimize_map` (src/Mapper.py:237-640) incl. its add_neural_points, frustum selection (`get_mask_from_c2w`), the
    geometry -> colour stage switch and the per-stage learning rates, five requested iterations;
  * `Renderer.render_img` (src/utils/Renderer.py:204-283) on a window camera, Replica (uniform zero-depth sampling) and TUM
    (sample_near_pcl) configurations, with zero-depth holes.
The harness only (a) records the `torch.randint` draws of src/common.py:66 so that the CUDA shells can be fed the same pixels,
(b) snapshots gradients inside `optimizer.step()`, (c) records the arguments / results of `render_batch_ray`, and (d) re-seeds
the RNG before every `render_batch_ray` of `render_img` so that every 3000-ray chunk draws the same two no-neighbour vectors
(the fused render draws them once per image).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H            # noqa: E402
from oracle.make_golden import fp16_exact, sparse_rows      # noqa: E402
from point_slam_b200 import synth              # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
INTR = synth.TUM_INTRINSICS
WIN = (150, 330, 220, 420)                     # pixel window (j0, j1, i0, i1) the golden scene's cloud was cropped to


def load_scene():
    z = np.load(os.path.join(OUT, 'scene.npz'))
    return dict(cloud=z['cloud'], geo_feats=z['geo_feats'].astype(np.float32), col_feats=z['col_feats'].astype(np.float32),
                c2w=z['c2w'])


def window_frame(c2w, seed, holes=0.0):
    """Synthetic frame of pose c2w, fp16-exact values, depth zero OUTSIDE the cloud's pixel window (those pixels are dropped by
    the callers' depth filter, common.py:173-179) and in a few holes inside it."""
    depth, color = synth.make_frame(c2w, INTR)
    r_add, r_query = synth.sobel_radius_map(color)
    j0, j1, i0, i1 = WIN
    d = np.zeros_like(depth)
    d[j0:j1, i0:i1] = fp16_exact(depth[j0:j1, i0:i1])
    if holes > 0:
        rng = np.random.default_rng(seed)
        d[rng.random(d.shape) < holes] = 0.0
    return dict(depth=d, color=fp16_exact(color), r_add=np.round(r_add, 4), r_query=np.round(r_query, 4), c2w=c2w)


def crop(a):
    j0, j1, i0, i1 = WIN
    return np.ascontiguousarray(a[j0:j1, i0:i1])


class Recorder:
    """Proxy for the `torch` name of a reference module: records randint draws, everything else passes through."""

    def __init__(self):
        self.draws = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def randint(self, *a, **kw):
        out = torch.randint(*a, **kw)
        self.draws.append(out.clone())
        return out


def recording_adam():
    class RecordingAdam(torch.optim.Adam):
        """torch.optim.Adam that snapshots every parameter's gradient when step() is entered."""

        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            self.snapshots = []

        def step(self, closure=None):
            self.snapshots.append([[None if p.grad is None else p.grad.detach().clone() for p in g['params']]
                                   for g in self.param_groups])
            return super().step(closure)
    return RecordingAdam


class TorchWithAdam:
    """`torch` as seen by src/Mapper.py: torch.optim.Adam is the recording subclass (the reference builds its optimizer inside
    optimize_map, Mapper.py:402)."""

    def __init__(self, adam_cls):
        self.optim = types.SimpleNamespace(Adam=self._make(adam_cls))
        self.made = []

    def _make(self, cls):
        def make(*a, **kw):
            o = cls(*a, **kw)
            self.made.append(o)
            return o
        return make

    def __getattr__(self, name):
        return getattr(torch, name)


def record_render(renderer):
    """Record results of every render_batch_ray and the two N(0, 0.01^2) no-neighbour vectors POINT.forward draws inside it
    (decoder.py:170-171, 387-388: the first RNG use of the call) by replaying the RNG state the call started from."""
    calls = []
    orig = renderer.render_batch_ray

    def wrapped(*a, **kw):
        state = torch.get_rng_state()
        ret = orig(*a, **kw)
        after = torch.get_rng_state()
        torch.set_rng_state(state)
        rg = torch.zeros([32]).normal_(mean=0, std=0.01)
        rc = torch.zeros([32]).normal_(mean=0, std=0.01)
        torch.set_rng_state(after)
        calls.append(dict(gt_depth=kw['gt_depth'].detach().clone(), depth=ret[0].detach().clone(), var=ret[1].detach().clone(),
                          color=ret[2].detach().clone(), valid=ret[3].clone(), n_rays=a[2].shape[0], rand_geo=rg, rand_col=rc))
        return ret
    renderer.render_batch_ray = wrapped
    return calls


@contextlib.contextmanager
def record_backward():
    """Values of the tensors `.backward()` is called on (the reference does not return its mapping loss)."""
    losses = []
    orig = torch.Tensor.backward

    def bw(self, *a, **k):
        losses.append(float(self.detach()))
        return orig(self, *a, **k)
    torch.Tensor.backward = bw
    try:
        yield losses
    finally:
        torch.Tensor.backward = orig


def run_tracker(scene):
    j0, j1, i0, i1 = WIN
    ov = {'tracking.ignore_edge_H': j0, 'tracking.ignore_edge_W': i0, 'tracking.pixels': 400}
    ref = H.load_reference('configs/Replica/room0.yaml', ov)
    cfg = ref['cfg']
    Tracker = H.load_tracker_class()
    rec = Recorder()
    ref['common'].torch = rec
    try:
        model = H.build_decoders(ref)
        npc = H.build_npc(ref, scene['cloud'], scene['geo_feats'], scene['col_feats'])
        renderer = H.build_renderer(ref, INTR, coef=cfg['rendering']['sigmoid_coef_tracker'])
        calls = record_render(renderer)
        fr = window_frame(scene['c2w'], 1, holes=0.02)
        from scipy.spatial.transform import Rotation
        rng = np.random.default_rng(4)
        q = np.roll(Rotation.from_matrix(scene['c2w'][:3, :3]).as_quat(), 1)
        cam0 = torch.tensor(np.concatenate([q + rng.normal(0, 2e-3, 4), scene['c2w'][:3, 3] + rng.normal(0, 5e-3, 3)]), dtype=torch.float32)
        quad = cam0[:4].clone().requires_grad_(True)
        T = cam0[4:].clone().requires_grad_(True)
        lr = cfg['tracking']['lr']
        opt = recording_adam()([{'params': [T], 'lr': lr}, {'params': [quad], 'lr': lr * 0.2}])      # Tracker.py:296-299
        me = types.SimpleNamespace(
            device='cpu', npc=npc, H=INTR['H'], W=INTR['W'], fx=INTR['fx'], fy=INTR['fy'], cx=INTR['cx'], cy=INTR['cy'],
            ignore_edge_W=cfg['tracking']['ignore_edge_W'], ignore_edge_H=cfg['tracking']['ignore_edge_H'],
            sample_with_color_grad=False, depth_limit=cfg['tracking']['depth_limit'], use_dynamic_radius=cfg['use_dynamic_radius'],
            dynamic_r_query=torch.from_numpy(fr['r_query']), renderer=renderer, decoders=model,
            npc_geo_feats=npc.get_geo_feats().detach().clone(), npc_col_feats=npc.get_col_feats().detach().clone(),
            cloud_pos=torch.tensor(npc.cloud_pos()).reshape(-1, 3), exposure_feat=None,
            handle_dynamic=cfg['tracking']['handle_dynamic'], use_color_in_tracking=cfg['tracking']['use_color_in_tracking'],
            w_color_loss=cfg['tracking']['w_color_loss'])
        gt_color, gt_depth = torch.from_numpy(fr['color']), torch.from_numpy(fr['depth'])
        out = dict(cam0=cam0.numpy(), lr=np.float32(lr), w_color=np.float32(me.w_color_loss), n_pixels=np.int64(cfg['tracking']['pixels']),
                   edge=np.array([j0, i0], np.int64), c2w=scene['c2w'], depth_win=crop(fr['depth']), color_win=crop(fr['color']),
                   r_query_win=crop(fr['r_query']))
        torch.manual_seed(100)
        for it in range(3):
            n_draw = len(rec.draws)
            cam = torch.cat([quad, T], 0)
            loss, closs, gloss = Tracker.optimize_cam_in_batch(me, cam, gt_color, gt_depth, cfg['tracking']['pixels'], opt)
            assert len(rec.draws) == n_draw + 1 and len(calls) == it + 1
            out[f'pix{it}'] = rec.draws[-1].numpy()
            out[f'rand_geo{it}'], out[f'rand_col{it}'] = calls[it]['rand_geo'].numpy(), calls[it]['rand_col'].numpy()
            out[f'depth{it}'], out[f'color{it}'], out[f'n_rays{it}'] = calls[it]['depth'].numpy(), calls[it]['color'].numpy(), np.int64(calls[it]['n_rays'])
            snap = opt.snapshots[-1]
            out[f'loss{it}'] = np.float32(loss)
            out[f'grad_T{it}'], out[f'grad_quad{it}'] = snap[0][0].numpy(), snap[1][0].numpy()
            out[f'cam_after{it}'] = torch.cat([quad, T], 0).detach().numpy().copy()
            print(f'tracker it {it}: loss {loss:.6f} |grad quad| {snap[1][0].abs().max():.4e} |grad T| {snap[0][0].abs().max():.4e}')
        np.savez_compressed(os.path.join(OUT, 'caller_tracker.npz'), **out)
    finally:
        ref['common'].torch = torch


def run_mapper(scene):
    ov = {'mapping.pixels': 6000, 'mapping.pixels_adding': 2000, 'mapping.pixels_based_on_color_grad': 0, 'mapping.mapping_window_size': 3,
          'mapping.keyframe_selection_method': 'global', 'mapping.save_selected_keyframes_info': False}
    ref = H.load_reference('configs/Replica/room0.yaml', ov)
    cfg = ref['cfg']
    Mapper, mapper_mod = H.load_mapper_class(return_module=True)
    rec = Recorder()
    ref['common'].torch = rec
    tw = TorchWithAdam(recording_adam())
    mapper_mod.torch = tw
    try:
        model = H.build_decoders(ref)
        npc = H.build_npc(ref, scene['cloud'], scene['geo_feats'], scene['col_feats'])
        renderer = H.build_renderer(ref, INTR, coef=cfg['rendering']['sigmoid_coef_mapper'])
        calls = record_render(renderer)
        # current frame = the scene pose, two keyframes = small motions of it (they see the same window of the cloud)
        poses = [scene['c2w'].copy() for _ in range(3)]
        poses[0][:3, 3] += np.array([0.04, 0.02, -0.01]); poses[1][:3, 3] += np.array([-0.03, 0.03, 0.02])
        frames = [window_frame(p, 10 + k, holes=0.02) for k, p in enumerate(poses)]
        t = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
        kf_dict = [dict(color=t(f['color']), depth=t(f['depth']), est_c2w=t(f['c2w']), gt_c2w=t(f['c2w']),
                        dynamic_r_query=t(f['r_query'], torch.float64)) for f in frames[:2]]
        cur = frames[2]
        me = types.SimpleNamespace(
            H=INTR['H'], W=INTR['W'], fx=INTR['fx'], fy=INTR['fy'], cx=INTR['cx'], cy=INTR['cy'], npc=npc, cfg=cfg, device='cpu',
            keyframe_selection_method='global', mapping_window_size=cfg['mapping']['mapping_window_size'], keyframe_dict=kf_dict,
            save_selected_keyframes_info=False, mapping_pixels=cfg['mapping']['pixels'], pixels_adding=cfg['mapping']['pixels_adding'],
            pixels_based_on_color_grad=0, use_dynamic_radius=cfg['use_dynamic_radius'], dynamic_r_add=t(cur['r_add'], torch.float64),
            dynamic_r_query=t(cur['r_query'], torch.float64), encode_exposure=False, frustum_feature_selection=True,
            frustum_edge=cfg['mapping']['frustum_edge'], fix_geo_decoder=cfg['mapping']['fix_geo_decoder'],
            fix_color_decoder=cfg['mapping']['fix_color_decoder'], decoders=model, BA=False, min_iter_ratio=cfg['mapping']['min_iter_ratio'],
            geo_iter_first=cfg['mapping']['geo_iter_first'], geo_iter_ratio=cfg['mapping']['geo_iter_ratio'], n_img=10 ** 6,
            color_refine=False, vis_inside=False, renderer=renderer, w_color_loss=cfg['mapping']['w_color_loss'], wandb=False,
            num_joint_iters=5, visualizer=types.SimpleNamespace(vis=lambda *a, **k: None), exposure_feat=None, gt_camera=False,
            save_rendered_image=False)
        me.get_mask_from_c2w = types.MethodType(Mapper.get_mask_from_c2w, me)
        n0 = npc.pts_num()
        dec0 = {k: v.detach().clone() for k, v in model.color_decoder.state_dict().items()}
        added = {}
        orig_add = npc.add_neural_points

        def add_and_snapshot(*a, **kw):
            ret = orig_add(*a, **kw)
            added['geo'], added['col'] = npc.get_geo_feats()[n0:].clone(), npc.get_col_feats()[n0:].clone()
            return ret
        npc.add_neural_points = add_and_snapshot
        torch.manual_seed(77)
        np.random.seed(77)
        with contextlib.redirect_stdout(io.StringIO()) as log, record_backward() as losses:
            Mapper.optimize_map(me, 5, torch.tensor(5), t(cur['color']), t(cur['depth']), t(cur['c2w']), kf_dict, [0, 1], t(cur['c2w']))
        print(log.getvalue().strip().splitlines()[0])
        opt = tw.made[-1]
        n_it = len(opt.snapshots)
        n1 = npc.pts_num()
        cp = np.asarray(npc.cloud_pos(), np.float32)
        out = dict(c2w=np.stack([f['c2w'] for f in frames]), depth_win=np.stack([crop(f['depth']) for f in frames]),
                   color_win=np.stack([crop(f['color']) for f in frames]), r_query_win=np.stack([crop(f['r_query']) for f in frames]),
                   r_add_win=crop(cur['r_add']), n_iters=np.int64(n_it), n_pixels=np.int64(cfg['mapping']['pixels']),
                   pix_add=rec.draws[0].numpy(), added_pos=cp[n0:n1], added_geo=added['geo'].numpy(), added_col=added['col'].numpy(),
                   w_color=np.float32(me.w_color_loss))
        assert len(rec.draws) == 1 + 3 * n_it and len(calls) == n_it and len(losses) == n_it and added['geo'].shape[0] == n1 - n0
        idx_mask = None
        stages = []
        for it in range(n_it):
            out[f'pix{it}'] = np.stack([rec.draws[1 + 3 * it + k].numpy() for k in range(3)])
            snap = opt.snapshots[it]
            stage = 'color' if snap[2][0] is not None else 'geometry'
            stages.append(stage)
            c = calls[it]
            m = (c['gt_depth'] > 0) & c['valid'] & (~torch.isnan(c['depth']))
            out[f'loss{it}'] = np.float32(losses[it])
            out[f'rand_geo{it}'], out[f'rand_col{it}'] = c['rand_geo'].numpy(), c['rand_col'].numpy()
            out[f'n_rays{it}'] = np.int64(c['n_rays'])
            out[f'n_valid{it}'] = np.int64(int(m.sum()))
            out[f'depth{it}'], out[f'color{it}'] = c['depth'].numpy(), c['color'].numpy()
            out[f'grad_geo_rows{it}'], out[f'grad_geo_vals{it}'] = sparse_rows(snap[1][0])        # rows of the (U,32) slice with a non-zero gradient
            if stage == 'color':
                out[f'grad_col_rows{it}'], out[f'grad_col_vals{it}'] = sparse_rows(snap[2][0])
                names = [k for k, _ in model.color_decoder.named_parameters()]
                for nm, g in zip(names, snap[0]):
                    if g is not None:
                        out[f'grad_dec{it}.{nm}'] = g.numpy()
            print(f'mapper it {it}: stage {stage} rays {c["n_rays"]} valid {int(m.sum())} loss {losses[it]:.6f}')
        out['stages'] = np.array(stages)
        # frustum selection the reference made (after the add) and the end state
        idx = np.asarray(me.get_mask_from_c2w(t(cur['c2w']), cur['depth']), np.int64)
        out['indices'] = idx
        # end state of the optimised rows (only those that moved: Adam leaves a row with an all-zero gradient history alone)
        g_after, c_after = npc.get_geo_feats().detach()[idx], npc.get_col_feats().detach()[idx]
        g_before = torch.cat([torch.from_numpy(scene['geo_feats']), added['geo']])[idx]
        c_before = torch.cat([torch.from_numpy(scene['col_feats']), added['col']])[idx]
        mg, mc = (g_after != g_before).any(1), (c_after != c_before).any(1)
        out['geo_after_rows'], out['geo_after_vals'] = torch.nonzero(mg)[:, 0].numpy(), g_after[mg].numpy()
        out['col_after_rows'], out['col_after_vals'] = torch.nonzero(mc)[:, 0].numpy(), c_after[mc].numpy()
        for k, v in model.color_decoder.state_dict().items():
            if not torch.equal(v, dec0[k]):
                out[f'dec_after.{k}'] = v.numpy()
        np.savez_compressed(os.path.join(OUT, 'caller_mapper.npz'), **out)
        print('mapper golden: iterations', n_it, stages, 'added points', n1 - n0, 'frustum rows', idx.shape[0])
    finally:
        ref['common'].torch = torch
        mapper_mod.torch = torch


def run_render_img(scene):
    """Renderer.render_img on the window camera (H=180, W=200, principal point shifted: pixel (i,j) == full-image pixel
    (i+220, j+150)), zero-depth holes; every 3000-ray chunk re-seeded."""
    j0, j1, i0, i1 = WIN
    intr = dict(H=j1 - j0, W=i1 - i0, fx=INTR['fx'], fy=INTR['fy'], cx=INTR['cx'] - i0, cy=INTR['cy'] - j0)
    for name, cfgfile in (('replica', 'configs/Replica/room0.yaml'), ('tum', 'configs/TUM_RGBD/freiburg1_desk.yaml')):
        ref = H.load_reference(cfgfile, {'cam.crop_edge': 0})
        model = H.build_decoders(ref)
        npc = H.build_npc(ref, scene['cloud'], scene['geo_feats'], scene['col_feats'])
        renderer = H.build_renderer(ref, intr)
        fr = window_frame(scene['c2w'], 31, holes=0.03)
        depth, r_query = crop(fr['depth']).copy(), crop(fr['r_query'])
        depth[40:60, 90:130] = 0.0                                  # a block of missing depth (incl. pixels looking into the cloud's hole)
        seed = 500
        orig = renderer.render_batch_ray

        def reseeded(*a, **kw):
            torch.manual_seed(seed)
            return orig(*a, **kw)
        renderer.render_batch_ray = reseeded
        c2w = torch.from_numpy(scene['c2w']).float()
        d, u, c = renderer.render_img(npc, model, c2w, 'cpu', 'color', gt_depth=torch.from_numpy(depth), npc_geo_feats=npc.get_geo_feats(),
                                      npc_col_feats=npc.get_col_feats(), dynamic_r_query=torch.from_numpy(r_query),
                                      cloud_pos=torch.tensor(npc.cloud_pos()).reshape(-1, 3))
        assert d.dtype == torch.float64 and u.dtype == torch.float64 and c.dtype == torch.float32
        ra, rb = H.draw_rand_vecs(seed)
        assert torch.equal(d.float().double(), d)
        np.savez_compressed(os.path.join(OUT, f'render_img_{name}.npz'), intr=np.array([intr[k] for k in ('H', 'W', 'fx', 'fy', 'cx', 'cy')]),
                            c2w=scene['c2w'], gt_depth=depth, r_query=r_query, rand_geo=ra.numpy(), rand_col=rb.numpy(),
                            depth=d.float().numpy(), uncertainty=u.float().numpy(), color=c.numpy())
        print(f'render_img {name}: depth [{float(d.min()):.3f}, {float(d.max()):.3f}] zero-depth px {int((depth == 0).sum())} '
              f'nan {int(torch.isnan(d).sum())}')


def main():
    torch.set_num_threads(8)
    scene = load_scene()
    what = sys.argv[1:] or ['tracker', 'mapper', 'render_img']
    if 'tracker' in what:
        run_tracker(scene)
    if 'mapper' in what:
        run_mapper(scene)
    if 'render_img' in what:
        run_render_img(scene)


if __name__ == '__main__':
    main()
