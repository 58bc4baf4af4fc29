"""CPU-only checks of the host side: the C-ABI library builds, loads and exports every symbol the header
declares; the drop-in modules keep the reference's state_dict keys and seeded initial values."""
import ctypes
import os
import re

import numpy as np
import torch

from point_slam_b200 import _lib
from point_slam_b200.default_config import make_cfg
from tests import cases as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    path = _lib.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, 'include', 'pointslam_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(psl_[a-z0-9_]+)\s*\(', header)))
    assert declared, 'no declarations found'
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert set(declared) == set(_lib.EXPORTS), 'ctypes signature table out of sync with the header'
    # calls that need no GPU
    _lib.load()
    assert _lib.load().psl_version() == 100
    assert _lib.load().psl_packed_params_floats() > 100000


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.DecoderParams) == 8 * _lib.N_PARAMS
    assert ctypes.sizeof(_lib.DecodeCfg) == 40
    assert ctypes.sizeof(_lib.Grid) == 48       # 3 pointers + u32 + i32 + 2 floats + the device meta pointer
    cfg = _lib.DecodeCfg(1, 1, 0, 0, 2, 5, 0, 0, 0.0064)
    assert _lib.load().psl_decode_save_floats_per_sample(ctypes.byref(cfg)) == 2304
    cfg.encode_rel_pos = 0
    assert _lib.load().psl_decode_save_floats_per_sample(ctypes.byref(cfg)) == 1024
    cfg.stage = 0
    assert _lib.load().psl_decode_save_floats_per_sample(ctypes.byref(cfg)) == 352


def test_seeded_module_matches_reference_init():
    """POINT(cfg) under manual_seed(1219) must reproduce the reference's seeded initial colour weights (golden
    decoders_base.npz was made by the reference constructor + pretrained geometry weights)."""
    from point_slam_b200.src.conv_onet import config as model_config
    gold = C.load_params(False)
    torch.manual_seed(1219)
    m = model_config.get_model(make_cfg('replica', 'cpu'))
    sd = m.state_dict()
    assert set(sd.keys()) == set(k for k in gold if k != 'color_decoder.embedder._B')
    for k, v in sd.items():
        if k.startswith('color_decoder.'):
            assert torch.equal(v, gold[k]), k
    assert torch.equal(m.color_decoder.embedder._B, gold['color_decoder.embedder._B'])
    assert 'embedder._B' not in m.color_decoder.state_dict()            # plain tensor, like the reference
    # exposure variant (ScanNet): extra mlp_exposure parameters, same keys as the reference
    gold_e = C.load_params(True)
    torch.manual_seed(1219)
    me = model_config.get_model(make_cfg('scannet', 'cpu'))
    assert set(me.state_dict().keys()) == set(k for k in gold_e if k != 'color_decoder.embedder._B')
    for k, v in me.state_dict().items():
        if k.startswith('color_decoder.'):
            assert torch.equal(v, gold_e[k]), k


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing somewhere else."""
    import pytest
    from point_slam_b200 import ops
    with pytest.raises(AssertionError):
        _lib.ptr(torch.zeros(4))
    g = ops.SpatialHash(0.08)
    with pytest.raises((AssertionError, RuntimeError)):
        g.build(torch.zeros(10, 3))


def test_add_neural_points_bookkeeping_with_stubbed_kernels(monkeypatch):
    """Host logic of NeuralPointCloud.add_neural_points (capacity-doubling buffers, adoption of caller-assigned tensors,
    lazy input lists, counts) with the two library calls replaced by the CPU oracle: the final state must equal what the
    unmodified reference produced (tests/golden/aux.npz).  The CUDA kernels themselves are checked on the GPU by
    tests/test_gpu_parity.py::test_add_neural_points_and_sample_near_pcl against the same vectors."""
    from oracle import point_slam_oracle as O
    from point_slam_b200 import ops
    from point_slam_b200.src.neural_point import NeuralPointCloud

    def fake_build(self, cloud_pos, appended_from=0):
        self._cloud = cloud_pos.detach().clone().float().reshape(-1, 3)
        self.n = self._cloud.shape[0]
        return self

    def fake_add_points(grid, ro, rd, dep, col, new_pos, radius, dynamic_radius=None, n_add=3, fixed_interval=False,
                        near_surface=0.98, far_surface=1.02):
        assert not fixed_interval
        m = dep > 0
        dyn = None
        if dynamic_radius is not None:
            dyn = torch.zeros(dep.shape[0], dtype=dynamic_radius.dtype)
            dyn[m] = dynamic_radius[:int(m.sum())]             # the kernel reads r2[rank among the depth > 0 rays]
        cloud = getattr(grid, '_cloud', None)
        keep, pts = O.add_points(cloud, ro, rd, dep, radius_add=radius, dynamic_radius=dyn, N_add=n_add,
                                 near_surface=near_surface, far_surface=far_surface)
        new_pos[:pts.shape[0]] = pts
        n = dep.shape[0]
        in_pos, in_rgb = torch.zeros(n, 3), torch.zeros(n, 3)
        k = int(keep.sum())
        in_pos[:k] = (ro[m] + rd[m] * dep[m][:, None])[keep]
        in_rgb[:k] = (col[m] * 255)[keep]
        return torch.tensor([int(m.sum()), k], dtype=torch.int32), in_pos, in_rgb

    monkeypatch.setattr(ops.SpatialHash, 'build', fake_build)
    monkeypatch.setattr(ops, 'add_points', fake_add_points)
    z = np.load(C.GOLDEN + '/aux.npz')
    scene = C.load_scene()
    npc = NeuralPointCloud(make_cfg('replica', 'cpu'))
    npc._cloud_pos = scene['cloud']                         # caller-assigned storage, like the offline tools do
    npc._pts_num = scene['cloud'].shape[0]
    npc.geo_feats, npc.col_feats = scene['geo_feats'].clone(), scene['col_feats'].clone()
    npc.index.add(npc._pos)
    ro, rd, gd = (torch.from_numpy(z[k]) for k in ('add_rays_o', 'add_rays_d', 'add_depth'))
    r_add = torch.from_numpy(z['add_r_add'])
    col = torch.rand(ro.shape[0], 3)
    n0 = npc.pts_num()
    k1 = npc.add_neural_points(ro, rd, gd, col, dynamic_radius=r_add[gd > 0])
    n1 = npc.pts_num()
    assert torch.equal(npc.get_geo_feats()[:n0], scene['geo_feats'])          # adopted, not lost
    k2 = npc.add_neural_points(ro, rd, gd, col, is_pts_grad=True)
    assert int(k1) == int(z['add_kept1']) and int(k2) == int(z['add_kept2'])
    cp = np.asarray(npc.cloud_pos(), dtype=np.float32)
    assert np.array_equal(cp[n0:n1], z['add_new1']) and np.array_equal(cp[n1:], z['add_new2'])
    assert npc.pts_num() == cp.shape[0] == npc.index_ntotal() == npc.get_geo_feats().shape[0] == npc.get_col_feats().shape[0]
    assert np.array_equal(np.asarray(npc.input_pos(), np.float32), z['add_input_pos'])
    assert len(npc.input_rgb()) == int(k1) + int(k2)
    assert npc._geo_buf.shape[0] >= npc.pts_num() and npc.get_geo_feats().data_ptr() == npc._geo_buf.data_ptr()
    new_rows = npc.get_geo_feats()[n0:]
    assert 0.05 < float(new_rows.std()) < 0.2                                  # fresh N(0, 0.1^2) rows
    import pytest
    with pytest.raises(AssertionError):                                        # the reference's shape assertion (:209)
        npc.add_neural_points(ro, rd, gd, col, dynamic_radius=r_add)


def test_checkpoint_format_matches_reference_logger(tmp_path):
    """Keys, key order, value types and file name of `Logger.log` equal what the unmodified reference Logger wrote for the
    same arguments in the build container (src/utils/Logger.py:20-44; the type table below was printed by that run)."""
    import types
    from point_slam_b200.src.utils.Logger import Logger, CKPT_KEYS
    ref_types = {'geo_feats': 'Tensor', 'col_feats': 'Tensor', 'cloud_pos': 'list', 'pts_num': 'int', 'input_pos': 'list',
                 'input_rgb': 'list', 'decoder_state_dict': 'OrderedDict', 'gt_c2w_list': 'Tensor', 'estimate_c2w_list': 'Tensor',
                 'keyframe_list': 'list', 'keyframe_dict': 'list', 'selected_keyframes': 'dict', 'idx': 'int',
                 'exposure_feat_all': 'Tensor'}
    dec = torch.nn.Linear(2, 2)
    m = types.SimpleNamespace(verbose=False, ckptsdir=str(tmp_path), gt_c2w_list=torch.zeros(3, 4, 4),
                              estimate_c2w_list=torch.zeros(3, 4, 4), decoders=dec)
    buf = torch.randn(16, 32)                                   # capacity buffer; the npc hands out a view of 5 rows
    npc = types.SimpleNamespace(get_geo_feats=lambda: buf[:5], get_col_feats=lambda: buf[:5] * 2, cloud_pos=lambda: [[0., 0., 0.]] * 5,
                                pts_num=lambda: 5, input_pos=lambda: [[0., 0., 0.]], input_rgb=lambda: [[1., 2., 3.]])
    path = Logger(None, None, m).log(7, [], [0], {}, npc, exposure_feat=[torch.zeros(1, 8)] * 2)
    assert os.path.basename(path) == '00007.tar' and os.path.exists(path)
    ck = torch.load(path, weights_only=False)
    assert list(ck.keys()) == list(ref_types.keys()) == list(CKPT_KEYS)
    assert {k: type(v).__name__ for k, v in ck.items()} == ref_types
    assert torch.equal(ck['geo_feats'], buf[:5]) and ck['geo_feats'].untyped_storage().nbytes() == 5 * 32 * 4   # exact size
    assert ck['exposure_feat_all'].shape == (2, 1, 8)
    path = Logger(None, None, m).log(8, [], [0], {}, npc)
    assert torch.load(path, weights_only=False)['exposure_feat_all'] is None


def test_replica_append_points_matches_source_cloud(monkeypatch):
    """parallel.make_delta / apply_delta on two real NeuralPointCloud objects (kernels stubbed by the oracle as above): after the
    mapping rank added points and changed feature rows, the replica brought up to date through `append_points` /
    `update_*_feats` holds bit-identical positions and features, and both keep valid row-prefix views of their buffers."""
    from oracle import point_slam_oracle as O
    from point_slam_b200 import ops, parallel as PL
    from point_slam_b200.src.neural_point import NeuralPointCloud

    def fake_build(self, cloud_pos, appended_from=0):
        self._cloud = cloud_pos.detach().clone().float().reshape(-1, 3)
        self.n = self._cloud.shape[0]
        return self

    def fake_add_points(grid, ro, rd, dep, col, new_pos, radius, dynamic_radius=None, n_add=3, fixed_interval=False,
                        near_surface=0.98, far_surface=1.02):
        m = dep > 0
        keep, pts = O.add_points(getattr(grid, '_cloud', None), ro, rd, dep, radius_add=radius, N_add=n_add,
                                 near_surface=near_surface, far_surface=far_surface)
        new_pos[:pts.shape[0]] = pts
        k = int(keep.sum())
        in_pos, in_rgb = torch.zeros(dep.shape[0], 3), torch.zeros(dep.shape[0], 3)
        in_pos[:k] = (ro[m] + rd[m] * dep[m][:, None])[keep]
        return torch.tensor([int(m.sum()), k], dtype=torch.int32), in_pos, in_rgb

    monkeypatch.setattr(ops.SpatialHash, 'build', fake_build)
    monkeypatch.setattr(ops, 'add_points', fake_add_points)
    z = np.load(C.GOLDEN + '/aux.npz')
    ro, rd, gd = (torch.from_numpy(z[k]) for k in ('add_rays_o', 'add_rays_d', 'add_depth'))
    dec = torch.nn.Module()
    dec.color_decoder = torch.nn.Linear(3, 2)
    src, rep = NeuralPointCloud(make_cfg('replica', 'cpu')), NeuralPointCloud(make_cfg('replica', 'cpu'))
    torch.manual_seed(4)
    src.add_neural_points(ro[:200], rd[:200], gd[:200], torch.zeros(200, 3))           # first batch: everything valid is kept
    d0 = PL.make_delta(src, dec, 0, None)
    PL.apply_delta(rep, dec, d0)
    for step in range(3):                                                              # grow past the first capacity
        n_before = src.pts_num()
        lo = 200 + 60 * step
        src.add_neural_points(ro[lo:lo + 60] + 5.0 * (step + 1), rd[lo:lo + 60], gd[lo:lo + 60], torch.zeros(60, 3))
        upd = torch.tensor([1, 5, n_before - 1])
        src.update_geo_feats(torch.full((3, 32), float(step)), upd)
        src.update_col_feats(torch.full((3, 32), -float(step)), upd)
        PL.apply_delta(rep, dec, PL.make_delta(src, dec, n_before, upd))
        assert rep.pts_num() == src.pts_num() == rep.index_ntotal() == src.index_ntotal() > n_before
        assert torch.equal(rep.cloud_pos_tensor(), src.cloud_pos_tensor())
        assert torch.equal(rep.get_geo_feats(), src.get_geo_feats()) and torch.equal(rep.get_col_feats(), src.get_col_feats())
        for npc in (src, rep):
            assert npc.get_geo_feats().data_ptr() == npc._geo_buf.data_ptr() and npc._geo_buf.shape[0] >= npc.pts_num()
            assert npc.cloud_pos_tensor().data_ptr() == npc._pos_buf.data_ptr()
    assert rep.cloud_pos() == src.cloud_pos()


def test_entry_points_reject_null_arguments_without_a_gpu():
    """Error behaviour of the C ABI (include/pointslam_b200.h: 0 or a negative code, message through psl_last_error, never a
    crash): NULL / malformed arguments are refused BEFORE any CUDA call, so this runs on the CPU-only build box."""
    lib = _lib.load()
    C_ = ctypes
    g = _lib.Grid(None, None, None, 0, 0, 0.08, 0.0)
    z12 = (C_.c_double * 12)()
    assert lib.psl_add_points(None, None, None, None, None, 10, None, 0.0016, 3, 0, 0.98, 1.02, None, None, None, None, None, None, 0, None) < 0
    assert b'grid' in lib.psl_last_error()
    assert lib.psl_add_points(C_.byref(g), None, None, None, None, 10, None, 0.0016, 3, 0, 0.98, 1.02, None, None, None, None, None, None, 0,
                              None) < 0
    assert b'NULL' in lib.psl_last_error()
    assert lib.psl_frustum_select(None, 10, z12, 500.0, 500.0, 320.0, 240.0, None, 480, 640, -4, None, None, None, None, 0, None) < 0
    assert b'NULL' in lib.psl_last_error()
    assert lib.psl_frustum_select(None, 10, None, 500.0, 500.0, 320.0, 240.0, None, 480, 640, -4, None, None, None, None, 0, None) < 0
    bad = _lib.Grid(None, None, None, 3, 5, 0.08, 0.0)                      # capacity not a power of two
    assert lib.psl_knn_query(C_.byref(bad), None, 0, None, 0.0064, 1, None, None, None, None) < 0
    assert b'capacity' in lib.psl_last_error()
    assert lib.psl_feat_scatter(None, 8, 100, None, None, None, None, None, None, None, 0, None) < 0
    assert lib.psl_composite_fwd(None, None, None, 4, 5, 0.1, None, None, None, None, None) < 0
    assert lib.psl_color_fwd_h2(None, None, None, 4, None, None, None, None, None, None, None, None, None, None, None) < 0
    assert lib.psl_add_points_ws_bytes(6000) > 4 * 6000 * 4 and lib.psl_frustum_select_ws_bytes(500000) > 500000 * 5
