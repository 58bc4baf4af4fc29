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
    assert ctypes.sizeof(_lib.Grid) == 40       # 3 pointers + u32 + i32 + 2 floats
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
