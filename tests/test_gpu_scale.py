"""GPU tests at BASELINE.json's full sizes through size-independent properties (the oracle would take minutes there),
full-image rendering, and the NCCL path when >= 2 GPUs are visible."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests import cases as C

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene(n_points, seed=1219, S=5, dataset='replica'):
    import types
    from point_slam_b200 import synth
    from point_slam_b200.default_config import make_cfg
    from point_slam_b200.src.conv_onet import config as model_config
    from point_slam_b200.src.neural_point import NeuralPointCloud
    from point_slam_b200.src.utils.Renderer import Renderer
    cfg = make_cfg(dataset, DEV, **{'rendering.N_surface': S})
    dec = model_config.get_model(cfg)
    P = C.load_params(False)
    emb = P.pop('color_decoder.embedder._B')
    dec.load_state_dict(P, strict=True)
    dec.color_decoder.embedder._B = emb
    dec = dec.to(DEV)
    cloud = synth.make_cloud(n_points, seed=seed)
    gf, cf = synth.make_features(n_points, seed=seed)
    npc = NeuralPointCloud(cfg)
    npc._cloud_pos = torch.from_numpy(cloud)
    npc._pts_num = n_points
    npc.geo_feats, npc.col_feats = torch.from_numpy(gf).to(DEV), torch.from_numpy(cf).to(DEV)
    npc.index.add(npc._pos)
    intr = synth.TUM_INTRINSICS
    ren = Renderer(cfg, None, types.SimpleNamespace(**{k: intr[k] for k in ('H', 'W', 'fx', 'fy', 'cx', 'cy')}))
    ren.sigmoid_coefficient = 0.1
    return cfg, dec, npc, ren, cloud


def test_c3_two_million_points_5k_rays_32_samples():
    """BASELINE config 3: 2 M points, 5000 rays x 32 samples.  Properties: kNN rows sorted by (D, index), all reported
    neighbours inside the radius, counts consistent, a 2000-query subset bit-exact against the kd-tree oracle; render
    finite, colour in [0,1], depth inside the sampled interval, deterministic across two calls."""
    from point_slam_b200 import ops, synth
    from oracle import point_slam_oracle as O
    from scipy.spatial import cKDTree
    S = 32
    cfg, dec, npc, ren, cloud = _scene(2_000_000, S=S)
    pose = synth.trajectory(3)[1]
    depth, color = synth.make_frame(pose)
    _, rq = synth.sobel_radius_map(color)
    o, d = synth.pixel_rays(pose, 480, 640, 517.3, 516.5, 318.6, 255.3)
    pix = np.random.default_rng(1).integers(0, 480 * 640, 5000)
    rays_o = torch.from_numpy(np.broadcast_to(o, (5000, 3)).astype(np.float32).copy()).to(DEV)
    rays_d = torch.from_numpy(d.reshape(-1, 3)[pix].astype(np.float32)).to(DEV)
    gd = torch.from_numpy(depth.reshape(-1)[pix]).to(DEV)
    dyn = torch.from_numpy(rq.reshape(-1)[pix]).to(DEV)
    with torch.no_grad():
        out1 = ren.render_batch_ray(npc, dec, rays_d, rays_o, DEV, 'color', gt_depth=gd, npc_geo_feats=npc.get_geo_feats(),
                                    npc_col_feats=npc.get_col_feats(), cloud_pos=npc.cloud_pos_tensor(), dynamic_r_query=dyn)
        torch.manual_seed(0)
        out2 = ren.render_batch_ray(npc, dec, rays_d, rays_o, DEV, 'color', gt_depth=gd, npc_geo_feats=npc.get_geo_feats(),
                                    npc_col_feats=npc.get_col_feats(), cloud_pos=npc.cloud_pos_tensor(), dynamic_r_query=dyn)
    dep, var, col, valid = out1
    assert torch.isfinite(dep).all() and torch.isfinite(var).all() and torch.isfinite(col).all()
    assert (col >= 0).all() and (col <= 1).all()
    v = valid
    assert float(v.float().mean()) > 0.9
    assert ((dep[v] >= 0.98 * gd[v] * (1 - 1e-5)) & (dep[v] <= 1.02 * gd[v] * (1 + 1e-5))).all()
    assert torch.equal(out1[0][v], out2[0][v]) and torch.equal(out1[3], out2[3])       # rays with neighbours do not depend on the random fill
    # kNN properties on the same sample points
    tv = torch.linspace(0., 1., S, device=DEV)
    z = 0.98 * gd[:, None] * (1. - tv) + 1.02 * gd[:, None] * tv
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]).reshape(-1, 3)
    D, I, n = ops.knn_query(npc.spatial_hash(), pts, dynamic_radius=dyn, group=S)
    r2 = (dyn.double() ** 2).repeat_interleave(S)
    found = I >= 0
    assert ((D.double() <= r2[:, None]) | ~found).all()
    Dm = torch.where(found, D, torch.full_like(D, float('inf')))
    assert (Dm[:, 1:] >= Dm[:, :-1]).all()
    tie = (Dm[:, 1:] == Dm[:, :-1]) & found[:, 1:]
    assert (I[:, 1:][tie] > I[:, :-1][tie]).all()
    assert torch.equal(n, ((D.double() < r2[:, None]) & found).sum(1).int())
    # every reported distance is the canonical distance to the reported point
    cp = npc.cloud_pos_tensor()
    sel = torch.randperm(pts.shape[0], device=DEV)[:2000]
    diff = cp[I[sel].clamp_min(0).long()] - pts[sel][:, None, :]
    Dre = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
    assert torch.equal(torch.where(found[sel], Dre, D[sel]), D[sel])
    tree = cKDTree(cloud.astype(np.float64))
    Do, Io, no = O.find_neighbors(torch.from_numpy(cloud), pts[sel].cpu(), 0.08, dyn.repeat_interleave(S)[sel].cpu(), tree=tree)
    inr = Do.double() <= r2[sel].cpu()[:, None]
    assert torch.equal(I[sel].cpu().long(), torch.where(inr, Io, torch.full_like(Io, -1)))
    assert torch.equal(n[sel].cpu(), no)


def test_render_img_single_call_matches_chunked_batches():
    """Renderer.render_img (one fused launch set for the image) == the reference's 3000-ray chunk loop of render_batch_ray."""
    from point_slam_b200 import synth
    cfg, dec, npc, ren, cloud = _scene(300_000, dataset='tum')
    ren.H, ren.W = 96, 128
    ren.fx, ren.fy, ren.cx, ren.cy = 103.46, 103.3, 63.72, 51.06
    pose = synth.trajectory(3)[1]
    intr = dict(H=96, W=128, fx=ren.fx, fy=ren.fy, cx=ren.cx, cy=ren.cy)
    depth, color = synth.make_frame(pose, intr, holes=0.05, seed=3)
    gd = torch.from_numpy(depth).to(DEV)
    dyn = torch.full((96, 128), 0.1, dtype=torch.float64, device=DEV)
    c2w = torch.from_numpy(pose[:3, :4].astype(np.float32)).to(DEV)
    fixed = (torch.zeros(32, device=DEV), torch.zeros(32, device=DEV))
    dec.draw_no_neighbor_vectors = lambda stage, device: fixed
    d, u, c = ren.render_img(npc, dec, c2w, DEV, 'color', gt_depth=gd, npc_geo_feats=npc.get_geo_feats(),
                             npc_col_feats=npc.get_col_feats(), dynamic_r_query=dyn, cloud_pos=npc.cloud_pos_tensor())
    assert d.dtype == torch.float64 and d.shape == (96, 128) and c.shape == (96, 128, 3)
    from point_slam_b200.src import common
    ro, rd = common.get_rays(96, 128, ren.fx, ren.fy, ren.cx, ren.cy, c2w, DEV)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    ds, cs = [], []
    with torch.no_grad():
        for i in range(0, 96 * 128, 3000):
            dd, _, cc, _ = ren.render_batch_ray(npc, dec, rd[i:i + 3000], ro[i:i + 3000], DEV, 'color', gt_depth=gd.reshape(-1)[i:i + 3000],
                                                npc_geo_feats=npc.get_geo_feats(), npc_col_feats=npc.get_col_feats(),
                                                cloud_pos=npc.cloud_pos_tensor(), dynamic_r_query=dyn.reshape(-1)[i:i + 3000])
            ds.append(dd); cs.append(cc)
    assert torch.equal(d.reshape(-1).float(), torch.cat(ds)) and torch.equal(c.reshape(-1, 3), torch.cat(cs))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_nccl_two_ranks_map_delta_and_bench():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests', 'multigpu_check.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert 'MULTIGPU OK' in res.stdout
