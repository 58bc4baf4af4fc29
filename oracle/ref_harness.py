"""Run the UNMODIFIED reference hot path on the CPU  --  TEST INFRASTRUCTURE ONLY.

Works only where /root/reference exists (the build container).  Nothing under
`tests/ -m gpu`, `smoke()` or `bench.py` imports this module.

What is stubbed and why:
  * `faiss` (faiss-gpu==1.7.2, env.yaml:96) is not installable here, so a fake
    `faiss` module provides `IndexIVFFlat`/`index_cpu_to_gpu`/... backed by the
    EXACT search of `oracle.point_slam_oracle.knn_exact`.  The reference class
    `src.neural_point.NeuralPointCloud` then runs unmodified on top of it.
  * `skimage` is only needed at import time of `src/common.py` (:6-7).
  * `load_mapper_class()` additionally stubs the import-time-only dependencies of src/Mapper.py (open3d, colorama,
    matplotlib, torchmetrics, pytorch_msssim) so that the unmodified `Mapper.get_mask_from_c2w` can be called unbound;
    cv2, numpy and scipy are the real packages.
  * two CPU-only breakages are patched: `quad2rotation`'s `.to(quad.get_device())`
    (src/common.py:238) and `POINT.forward('geometry')`'s `device='cuda:-1'`
    (src/conv_onet/models/decoder.py:499,505).
"""
from __future__ import annotations

import os
import sys
import types
from contextlib import contextmanager

import numpy as np
import torch

REF_ROOT = '/root/reference'
_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))

from oracle import point_slam_oracle as O  # noqa: E402


class _ExactIndex:
    """Duck-type of the faiss index the reference builds at src/neural_point.py:37-41."""

    def __init__(self):
        self.is_trained = False
        self.nprobe = 1
        self._pts = np.zeros((0, 3), np.float32)

    @property
    def ntotal(self):
        return self._pts.shape[0]

    def train(self, x):
        self.is_trained = True

    def add(self, x):
        x = x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
        self._pts = np.concatenate([self._pts, x.astype(np.float32).reshape(-1, 3)], 0)

    def search(self, q, k):
        D, I = O.knn_exact(self._pts, q, k)
        return D, I


def _install_stubs():
    if 'faiss' in sys.modules and getattr(sys.modules['faiss'], '_psl_stub', False):
        return
    faiss = types.ModuleType('faiss')
    faiss._psl_stub = True
    faiss.METRIC_L2 = 1
    faiss.StandardGpuResources = lambda: object()
    faiss.IndexFlatL2 = lambda d: ('flat', d)
    faiss.IndexIVFFlat = lambda quantizer, d, nlist, metric: _ExactIndex()
    faiss.index_cpu_to_gpu = lambda res, dev, index: index
    contrib = types.ModuleType('faiss.contrib')
    tu = types.ModuleType('faiss.contrib.torch_utils')
    faiss.contrib = contrib
    contrib.torch_utils = tu
    sys.modules.update({'faiss': faiss, 'faiss.contrib': contrib, 'faiss.contrib.torch_utils': tu})
    sk = types.ModuleType('skimage')
    skc = types.ModuleType('skimage.color')
    skf = types.ModuleType('skimage.filters')
    skc.rgb2gray = lambda x: x
    sk.color, sk.filters = skc, skf
    sys.modules.update({'skimage': sk, 'skimage.color': skc, 'skimage.filters': skf})


class _TorchProxy:
    """`torch` as seen by the reference decoder module, with the 'cuda:-1' device string mapped to cpu."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def zeros(*a, **kw):
        if str(kw.get('device', '')).startswith('cuda:-1'):
            kw['device'] = 'cpu'
        return torch.zeros(*a, **kw)


@contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load_reference(config='configs/Replica/room0.yaml', overrides=None):
    """-> dict(cfg, modules...) with the reference imported from REF_ROOT."""
    assert os.path.isdir(REF_ROOT), 'reference tree not present (this harness only runs in the build container)'
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with _cwd(REF_ROOT):
        from src import config as ref_config
        import src.common as ref_common
        import src.neural_point as ref_np
        import src.utils.Renderer as ref_renderer
        import src.conv_onet.models.decoder as ref_decoder
        cfg = ref_config.load_config(config, 'configs/point_slam.yaml')
    cfg['mapping']['device'] = 'cpu'
    cfg['tracking']['device'] = 'cpu'
    for path, val in (overrides or {}).items():
        node = cfg
        keys = path.split('.')
        for k in keys[:-1]:
            node = node[k]
        node[keys[-1]] = val
    ref_decoder.torch = _TorchProxy()
    ref_common.quad2rotation = O.quad_to_rotation      # same 9 formulas, minus the CPU-breaking device line
    return dict(cfg=cfg, common=ref_common, neural_point=ref_np, renderer=ref_renderer, decoder=ref_decoder)


def build_decoders(ref, seed=1219, pretrained=True):
    """POINT(cfg) under a fixed seed + the pretrained geometry weights (Point_SLAM.py:143-164)."""
    torch.manual_seed(seed)
    cfg = ref['cfg']
    m = ref['decoder'].POINT(cfg, c_dim=cfg['model']['c_dim'], pos_embedding_method='fourier',
                             use_view_direction=cfg['model']['use_view_direction'])
    if pretrained:
        ck = torch.load(os.path.join(REF_ROOT, 'pretrained/middle_fine.pt'), weights_only=False, map_location='cpu')
        sd = {k[8 + 7:]: v for k, v in ck['model'].items() if k.startswith('coarse.decoder.')}
        m.geo_decoder.load_state_dict(sd, strict=False)
    return m


def decoder_params(model):
    """Flat dict for the oracle: state_dict + the non-registered colour embedder matrix."""
    P = {k: v.detach().clone() for k, v in model.state_dict().items()}
    P['color_decoder.embedder._B'] = model.color_decoder.embedder._B.detach().clone()
    return P


def build_npc(ref, cloud_pos, geo_feats, col_feats):
    """A real reference NeuralPointCloud (exact-search index) holding the given cloud."""
    npc = ref['neural_point'].NeuralPointCloud(ref['cfg'])
    pts = torch.as_tensor(cloud_pos, dtype=torch.float32)
    npc._cloud_pos = pts.tolist()
    npc._pts_num = pts.shape[0]
    npc.geo_feats = torch.as_tensor(geo_feats).clone()
    npc.col_feats = torch.as_tensor(col_feats).clone()
    npc.index.train(pts)
    npc.index.add(pts)
    return npc


def build_renderer(ref, intr, coef=0.1):
    slam = types.SimpleNamespace(**{k: intr[k] for k in ('H', 'W', 'fx', 'fy', 'cx', 'cy')})
    r = ref['renderer'].Renderer(ref['cfg'], None, slam)
    r.sigmoid_coefficient = coef
    return r


def draw_rand_vecs(seed):
    """The two N(0,0.01^2) no-neighbour vectors the reference draws per POINT.forward('color')
    (decoder.py:170-171 then :387-388) after `torch.manual_seed(seed)`."""
    torch.manual_seed(seed)
    a = torch.zeros([32]).normal_(mean=0, std=0.01)
    b = torch.zeros([32]).normal_(mean=0, std=0.01)
    return a, b


def _stub_caller_imports():
    """Import-time-only dependencies of src/Tracker.py / src/Mapper.py that are absent here (never called by the methods the
    goldens exercise); cv2, numpy, scipy are the real packages."""
    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    stub('open3d')
    stub('colorama', Fore=types.SimpleNamespace(), Style=types.SimpleNamespace())
    stub('matplotlib')
    stub('matplotlib.pyplot')
    stub('torchmetrics')
    stub('torchmetrics.image')
    stub('torchmetrics.image.lpip', LearnedPerceptualImagePatchSimilarity=object)
    stub('pytorch_msssim', ms_ssim=None)
    stub('wandb')
    sk = sys.modules['skimage']
    if not hasattr(sk.filters, 'sobel_h'):
        sk.filters.sobel_h = sk.filters.sobel_v = None


def load_mapper_class(return_module=False):
    """The reference `Mapper` class (src/Mapper.py), for its unbound methods (get_mask_from_c2w, optimize_map)."""
    assert os.path.isdir(REF_ROOT), 'reference tree not present (this harness only runs in the build container)'
    _install_stubs()
    _stub_caller_imports()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with _cwd(REF_ROOT):
        import src.Mapper as ref_mapper
    return (ref_mapper.Mapper, ref_mapper) if return_module else ref_mapper.Mapper


def load_tracker_class():
    """The reference `Tracker` class (src/Tracker.py), for its unbound method optimize_cam_in_batch."""
    assert os.path.isdir(REF_ROOT), 'reference tree not present (this harness only runs in the build container)'
    _install_stubs()
    _stub_caller_imports()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with _cwd(REF_ROOT):
        import src.Tracker as ref_tracker
    return ref_tracker.Tracker
