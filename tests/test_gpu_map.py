"""GPU tests of the map-maintenance kernels: psl_add_points and psl_frustum_select through
the drop-in API, against vectors frozen from the unmodified reference (tests/golden/aux.npz, frustum.npz), the CPU oracle
at full size, and torch restatements of the reference's op sequence for the configuration branches the golden scene does
not take.  Index / integer / position outputs are compared bit for bit."""
import numpy as np
import pytest
import torch

from oracle import point_slam_oracle as O
from point_slam_b200 import synth
from point_slam_b200.default_config import make_cfg
from tests import cases as C

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_frustum_select_matches_reference_golden():
    from point_slam_b200 import ops
    z = np.load(C.GOLDEN + '/frustum.npz')
    H, W, fx, fy, cx, cy = z['intr']
    cloud = torch.from_numpy(z['cloud']).to(DEV)
    for k in range(2):
        idx, mask = ops.frustum_select(cloud, torch.from_numpy(z[f'c2w{k}']), torch.from_numpy(z[f'depth{k}']).to(DEV), int(H), int(W),
                                       fx, fy, cx, cy, edge=int(z[f'edge{k}']), return_mask=True)
        assert idx.dtype == torch.int64 and np.array_equal(idx.cpu().numpy(), z[f'indices{k}']), k
        assert np.array_equal(torch.nonzero(mask).reshape(-1).cpu().numpy(), z[f'indices{k}'])
    # (3,4) pose, numpy depth and an empty cloud are accepted
    idx = ops.frustum_select(cloud, z['c2w0'][:3], z['depth0'], int(H), int(W), fx, fy, cx, cy, edge=int(z['edge0']))
    assert np.array_equal(idx.cpu().numpy(), z['indices0'])
    assert ops.frustum_select(cloud[:0], z['c2w0'], z['depth0'], int(H), int(W), fx, fy, cx, cy).numel() == 0
    # reuse=True: same indices out of the per-device scratch, call after call (the view is overwritten by the next call)
    for k in (0, 1, 0):
        idx = ops.frustum_select(cloud, z[f'c2w{k}'], z[f'depth{k}'], int(H), int(W), fx, fy, cx, cy, edge=int(z[f'edge{k}']), reuse=True)
        assert np.array_equal(idx.cpu().numpy(), z[f'indices{k}']), k


def test_frustum_select_full_size_vs_oracle():
    """BASELINE size (500k points, 640x480 depth with holes): identical index list as the CPU oracle, which is pinned to
    the reference on the golden scene."""
    from point_slam_b200 import ops
    intr = synth.TUM_INTRINSICS
    cloud = synth.make_cloud(500000, seed=3)
    for k, c2w in enumerate(synth.trajectory(9, seed=11)[[1, 6]]):
        depth, _ = synth.make_frame(c2w, intr, noise=True, holes=0.05, seed=k)
        c2w32 = c2w.astype(np.float32)
        want = O.frustum_indices(cloud, c2w32, depth, intr['H'], intr['W'], intr['fx'], intr['fy'], intr['cx'], intr['cy'], edge=-4)
        got = ops.frustum_select(torch.from_numpy(cloud).to(DEV), c2w32, torch.from_numpy(depth).to(DEV), intr['H'], intr['W'],
                                 intr['fx'], intr['fy'], intr['cx'], intr['cy'], edge=-4)
        assert 10000 < want.size < 400000
        assert np.array_equal(got.cpu().numpy(), want), k


def _torch_add_reference(npc_points, ro, rd, gd, radius, n_add, fixed, near, far):
    """The op sequence of neural_point.py:107-145 in torch on the device (brute-force neighbour test, float32 D)."""
    m = gd > 0
    o, d, dep = ro[m], rd[m], gd[m]
    pts_gt = o + d * dep[:, None]
    keep = torch.ones(pts_gt.shape[0], dtype=torch.bool, device=ro.device)
    if npc_points is not None and npc_points.shape[0]:
        diff = npc_points[None, :, :] - pts_gt[:, None, :]
        D = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
        keep = (D < radius ** 2).sum(-1) == 0
    dep_rep = dep.unsqueeze(-1).repeat(1, n_add)
    if fixed:
        zv = dep_rep + torch.linspace(-0.04, 0.04, steps=n_add).to(ro.device).unsqueeze(0)
    else:
        t = torch.linspace(0.0, 1.0, steps=n_add).to(ro.device)
        zv = near * dep_rep * (1. - t) + far * dep_rep * t
    pts = (o[..., None, :] + d[..., None, :] * zv[..., :, None])[keep].reshape(-1, 3)
    return keep, pts, pts_gt[keep]


@pytest.mark.parametrize('fixed,n_add', [(False, 3), (True, 3), (False, 5), (True, 4)])
def test_add_points_branches_vs_torch_sequence(fixed, n_add):
    """First call on an empty cloud (every depth > 0 ray kept), then a second batch filtered against the first, for the
    fixed-interval branch (neural_point.py:131-133) and other N_add -- bit-identical positions, order and lists."""
    from point_slam_b200.src.neural_point import NeuralPointCloud
    from point_slam_b200.src import common
    cfg = make_cfg('replica', DEV)
    cfg['pointcloud']['fix_interval_when_add_along_ray'] = fixed
    cfg['pointcloud']['N_add'] = n_add
    npc = NeuralPointCloud(cfg)
    intr = synth.TUM_INTRINSICS
    c2w = synth.look_at([1.6, 2.6, 1.1], [1.5, 0.0, 0.7])
    depth, color = synth.make_frame(c2w, intr, holes=0.1)
    g = torch.Generator().manual_seed(5)
    c2w_t = torch.from_numpy(c2w[:3, :4]).float().to(DEV)
    have = None
    for call in range(2):
        jj = torch.randint(100, 380, (700,), generator=g)
        ii = torch.randint(100, 540, (700,), generator=g)
        ro, rd = common.get_rays_from_uv(ii.float().to(DEV), jj.float().to(DEV), c2w_t, intr['fx'], intr['fy'], intr['cx'], intr['cy'], DEV)
        gd = torch.from_numpy(depth[jj.numpy(), ii.numpy()]).to(DEV)
        gc = torch.from_numpy(color[jj.numpy(), ii.numpy()]).to(DEV)
        radius = cfg['pointcloud']['radius_add']
        keep, pts, pts_gt = _torch_add_reference(have, ro, rd, gd, radius, n_add, fixed,
                                                 cfg['pointcloud']['near_end_surface'], cfg['pointcloud']['far_end_surface'])
        n0 = npc.pts_num()
        k = npc.add_neural_points(ro, rd, gd, gc)
        assert int(k) == int(keep.sum()) and (call == 0 or 0 < int(k) < int((gd > 0).sum()))
        assert torch.equal(npc.cloud_pos_tensor()[n0:], pts)
        assert npc.pts_num() == n0 + pts.shape[0] == npc.index_ntotal() == npc.get_col_feats().shape[0]
        have = npc.cloud_pos_tensor().clone()
    assert len(npc.input_pos()) == len(npc.input_rgb()) == npc.pts_num() // n_add
    assert np.array_equal(np.asarray(npc.input_pos()[-pts_gt.shape[0]:], np.float32), pts_gt.cpu().numpy())
    assert np.allclose(np.asarray(npc.input_rgb()[-pts_gt.shape[0]:], np.float32), (gc[gd > 0][keep] * 255).cpu().numpy(), rtol=0, atol=0)
    # renders keep working on the grown cloud: the kNN of a new point finds itself first
    D, I, nn = npc.find_neighbors_faiss(npc.cloud_pos_tensor()[-5:], step='query')
    assert torch.equal(I[:, 0].cpu(), torch.arange(npc.pts_num() - 5, npc.pts_num()))
    assert float(D[:, 0].max()) == 0.0 and int(nn.min()) >= 1


def test_add_points_many_appends_keep_features_and_capacity():
    """Amortised growth: rows written earlier survive re-allocations; capacity at least doubles when it grows."""
    from point_slam_b200.src.neural_point import NeuralPointCloud
    npc = NeuralPointCloud(make_cfg('replica', DEV))
    g = torch.Generator(device=DEV).manual_seed(1)
    caps, first_rows = [], None
    for it in range(6):
        n = 1500
        ro = torch.rand(n, 3, device=DEV, generator=g) * 4 + torch.tensor([10.0 * it, 0, 0], device=DEV)   # disjoint regions
        rd = torch.nn.functional.normalize(torch.randn(n, 3, device=DEV, generator=g), dim=1)
        gd = torch.rand(n, device=DEV, generator=g) + 0.5
        k = npc.add_neural_points(ro, rd, gd, torch.zeros(n, 3, device=DEV))
        assert 0 < int(k) <= n
        if first_rows is None:
            first_rows = (npc.get_geo_feats()[:100].clone(), npc.get_col_feats()[:100].clone(), npc.cloud_pos_tensor()[:100].clone())
        caps.append(npc._geo_buf.shape[0])
    assert torch.equal(npc.get_geo_feats()[:100], first_rows[0]) and torch.equal(npc.get_col_feats()[:100], first_rows[1])
    assert torch.equal(npc.cloud_pos_tensor()[:100], first_rows[2])
    grown = [b / a for a, b in zip(caps, caps[1:]) if b != a]
    assert grown and min(grown) >= 2.0 and len(grown) < 5
    assert npc.pts_num() == npc.index_ntotal() == npc.get_geo_feats().shape[0]


def test_checkpoint_round_trip_restores_the_cloud(tmp_path):
    """Logger.log -> torch.load -> load_neural_point_cloud into a fresh NeuralPointCloud (the restore sequence of
    get_mesh_tsdf_fusion.py:64-82): same positions, features, lists and bit-identical kNN answers."""
    import types
    from tests.gpu_harness import build_objects
    from point_slam_b200.src.neural_point import NeuralPointCloud
    from point_slam_b200.src.utils.Logger import Logger, load_neural_point_cloud
    z = np.load(C.GOLDEN + '/aux.npz')
    c = C.load_case('mapper_color')
    cfg, decoders, npc, renderer = build_objects(c)
    ro, rd, gd = (torch.from_numpy(z[k]).to(DEV) for k in ('add_rays_o', 'add_rays_d', 'add_depth'))
    npc.add_neural_points(ro, rd, gd, torch.rand(ro.shape[0], 3, device=DEV), is_pts_grad=True)     # grown cloud: views of buffers
    m = types.SimpleNamespace(verbose=False, ckptsdir=str(tmp_path), gt_c2w_list=torch.zeros(2, 4, 4),
                              estimate_c2w_list=torch.zeros(2, 4, 4), decoders=decoders)
    path = Logger(cfg, None, m).log(3, [], [0], None, npc)
    ck = torch.load(path, weights_only=False, map_location='cpu')
    assert ck['geo_feats'].shape[0] == ck['pts_num'] == len(ck['cloud_pos']) == npc.pts_num()
    assert ck['geo_feats'].untyped_storage().nbytes() == npc.pts_num() * 32 * 4
    npc2 = NeuralPointCloud(cfg)
    assert load_neural_point_cloud(npc2, ck, DEV) == npc.pts_num()
    assert torch.equal(npc2.cloud_pos_tensor(), npc.cloud_pos_tensor())
    assert torch.equal(npc2.get_geo_feats(), npc.get_geo_feats()) and torch.equal(npc2.get_col_feats(), npc.get_col_feats())
    assert npc2.input_pos() == npc.input_pos() and npc2.input_rgb() == npc.input_rgb()
    q = torch.from_numpy(z['knn_pts']).to(DEV)
    for a, b in zip(npc.find_neighbors_faiss(q, step='query'), npc2.find_neighbors_faiss(q, step='query')):
        assert torch.equal(a, b)
    decoders.load_state_dict(ck['decoder_state_dict'])
