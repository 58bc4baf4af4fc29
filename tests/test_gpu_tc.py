"""tcgen05 building blocks (psl_tc.cuh): 3xTF32 GEMM with fp32 TMEM accumulation must match fp32/fp64 matmul."""
import os

import pytest
import torch

from tests import cases as C

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('K,N', [(8, 16), (40, 128), (56, 128), (128, 32), (160, 128)])
def test_tc_gemm_3xtf32(mode, K, N):
    from point_slam_b200 import _lib as L
    lib = L.load_test_lib()
    if mode == 1 and K > 64:
        pytest.skip('SS-form test operand does not fit in shared memory')
    g = torch.Generator().manual_seed(K * 1000 + N + mode)
    A = (torch.randn(128, K, generator=g) * 3).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    D = torch.zeros(128, N, device='cuda')
    scratch = torch.empty(2 * N * K, device='cuda')
    L.check(lib.psl_tc_gemm_test(L.ptr(A), L.ptr(W), L.ptr(D), L.ptr(scratch), K, N, mode, L.stream()), 'psl_tc_gemm_test')
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t())
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    err32 = float(((A @ W.t()).double() - ref).abs().max() / ref.abs().max())
    print(f'mode {mode} K {K} N {N}: 3xTF32 err {err:.2e}  (fp32 matmul err {err32:.2e})')
    assert err < 2e-6


@pytest.mark.parametrize('name', ['mapper_color', 'tum_near_pcl', 'exposure_tracker', 's32_color', 'fixed_radius_zero_depth'])
def test_color_branch_on_tensor_cores_matches_oracle(name):
    """Inference render with the tcgen05 colour branch (3xTF32, TMEM-resident activations) vs the fp64 oracle and vs
    the fp32 FFMA kernel, on the golden cases (all colour configurations: rel-pos on/off, exposure affine, S = 32)."""
    import numpy as np
    from point_slam_b200 import ops
    from tests import cases as C
    from tests.gpu_harness import build_objects, DEV
    c = C.load_case(name)
    o64 = C.run_oracle(c, torch.float64)
    o32 = C.run_oracle(c, torch.float32)
    cfg, decoders, npc, renderer = build_objects(c)
    t = lambda k, dt=torch.float32: torch.from_numpy(np.asarray(c[k])).to(device=DEV, dtype=dt)
    rg, rc = t('rand_geo'), t('rand_col')
    decoders.draw_no_neighbor_vectors = lambda stage, device: (rg, rc if stage == 'color' else None)
    if 'cam_tensor' in c:
        from point_slam_b200.src import common
        c2w = common.get_camera_from_tensor(t('cam_tensor'))
        rays_o, rays_d = common.get_rays_from_uv(t('pix_i'), t('pix_j'), c2w, C.INTR['fx'], C.INTR['fy'], C.INTR['cx'], C.INTR['cy'], DEV)
    else:
        rays_o, rays_d = t('rays_o'), t('rays_d')
    dyn = t('dynamic_r_query', torch.float64) if 'dynamic_r_query' in c else None
    ef = t('exposure_feat') if 'exposure_feat' in c else None
    outs = {}
    for use_tc in (True, False):
        ops.USE_TENSOR_CORES = use_tc
        with torch.no_grad():
            outs[use_tc] = renderer.render_batch_ray(npc, decoders, rays_d, rays_o, DEV, c['stage'], gt_depth=t('gt_depth'),
                                                     npc_geo_feats=npc.get_geo_feats(), npc_col_feats=npc.get_col_feats(),
                                                     is_tracker=c['is_tracker'], cloud_pos=npc.cloud_pos_tensor(),
                                                     dynamic_r_query=dyn, exposure_feat=ef)
    ops.USE_TENSOR_CORES = True
    d_tc, v_tc, c_tc, m_tc = outs[True]
    d_ff, v_ff, c_ff, m_ff = outs[False]
    assert torch.equal(m_tc, m_ff)
    assert C.rel_err(d_tc.cpu(), d_ff.cpu()) < 1e-5                            # geometry branch: mma.sync 3xTF32 kernel vs the FFMA kernel
    assert not torch.equal(c_tc, c_ff), 'tensor-core path was not taken'
    assert C.rel_err(c_tc.cpu(), c_ff.cpu()) < 2e-5
    e_tc = C.rel_err(c_tc.cpu(), o64['color'])
    e_ff = C.rel_err(c_ff.cpu(), o64['color'])
    e_32 = C.rel_err(o32['color'], o64['color'])
    print(f'{name}: colour rel err vs fp64  tcgen05 {e_tc:.2e}  ffma {e_ff:.2e}  fp32 oracle {e_32:.2e}')
    assert e_tc <= max(1e-4, 3 * e_32)


@pytest.mark.parametrize('name', ['tracker_color', 'tum_tracker', 'mapper_color', 'tum_near_pcl', 's32_color', 'fixed_radius_zero_depth'])
def test_tensor_core_backward_data_path(name):
    """With the decoder frozen (what the tracker needs) the colour branch runs forward AND backward on tcgen05
    (psl_color_fwd_tc -> psl_color_bwd_tc): pose and feature gradients against the fp64 oracle."""
    from point_slam_b200 import ops
    from tests import cases as C
    from tests.gpu_harness import run_case_gpu
    c = C.load_case(name)
    o32, o64 = C.run_oracle(c, torch.float32), C.run_oracle(c, torch.float64)
    assert ops.USE_TENSOR_CORES and ops.USE_TC_BACKWARD
    got = run_case_gpu(c, freeze_decoders=True)
    ops.USE_TC_BACKWARD = False
    ref = run_case_gpu(c, freeze_decoders=True)                     # same forward, FFMA backward
    ops.USE_TC_BACKWARD = True

    def tol(what, g, a32, a64):
        e, e32 = C.rel_err(g.cpu(), a64), C.rel_err(a32, a64)
        print(f'{name} {what}: tcgen05 {e:.2e}  fp32 oracle {e32:.2e}')
        assert e <= max(1e-4, 3 * e32), what
    tol('loss', got['loss'], o32['loss'], o64['loss'])
    if 'grad_cam' in o64:
        tol('pose grad', got['grad_cam'], o32['grad_cam'], o64['grad_cam'])
        assert not torch.equal(got['grad_cam'], ref['grad_cam']), 'tensor-core backward was not taken'
    tol('geo feature grad', got['grad_geo'], o32['grad_geo'], o64['grad_geo'])
    tol('col feature grad', got['grad_col'], o32['grad_col'], o64['grad_col'])
    assert not got['grad_params']


@pytest.mark.parametrize('mode', [2, 3])
@pytest.mark.parametrize('K,N', [(16, 16), (48, 128), (64, 128), (128, 32), (128, 128), (208, 128)])
def test_tc_gemm_f16_planes(mode, K, N):
    """kind::f16 MMAs on 16-bit operand planes (hi = f16(a), lo = f16(a - hi); TS form keeps two k per TMEM column): the three
    products hi*hi + lo*hi + hi*lo must reproduce the fp32 product as well as 3xTF32 does, at half the TMEM / shared-memory
    footprint and twice the MMA rate.  (Mixing f16 and bf16 operands in one instruction faults on sm_100a:
    profiles/r02_s2_f16_planes_probe.log.)"""
    from point_slam_b200 import _lib as L
    lib = L.load_test_lib()
    if mode == 2 and K > 128:
        pytest.skip('TS-form A planes hold K <= 128')
    g = torch.Generator().manual_seed(K * 1000 + N + mode)
    A = (torch.randn(128, K, generator=g) * 3).cuda()
    A[:, 0] *= 1e-4                       # small magnitudes: both planes go subnormal (absolute error <= 3e-8)
    W = torch.randn(N, K, generator=g).cuda()
    ref = (A.double() @ W.double().t())
    err32 = float(((A @ W.t()).double() - ref).abs().max() / ref.abs().max())
    D = torch.zeros(128, N, device='cuda')
    scratch = torch.empty(N * K, device='cuda')
    L.check(lib.psl_tc_gemm_test_h(L.ptr(A), L.ptr(W), L.ptr(D), L.ptr(scratch), K, N, mode, 2, L.stream()), 'psl_tc_gemm_test_h')
    torch.cuda.synchronize()
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    print(f'mode {mode} K {K} N {N}: f16 hi/lo planes err {err:.2e} (fp32 matmul err {err32:.2e})')
    assert err < 2e-6


@pytest.mark.parametrize('name', ['mapper_color', 'tracker_color', 'tum_tracker'])
def test_f16_plane_kernels_agree_with_3xtf32_kernels(name):
    """The f16 hi/lo-plane forward / backward (psl_color_h2.cu, psl_color_bwd_h2.cu) against the 3xTF32 kernels they replace
    (psl_color_tc_w16.cu, psl_color_bwd_tc_w16.cu; PSL_H2=0 / PSL_H2_BWD=0): two error-compensated evaluations of the same
    fp32 mathematics -- outputs and every gradient agree far below the fp32 noise floor of the reference."""
    from point_slam_b200 import ops
    from tests import cases as C
    from tests.gpu_harness import run_case_gpu
    c = C.load_case(name)
    got = run_case_gpu(c)
    ops.USE_H2_FORWARD = ops.USE_H2_BACKWARD = False
    try:
        ref = run_case_gpu(c)
    finally:
        ops.USE_H2_FORWARD = ops.USE_H2_BACKWARD = True
    assert not torch.equal(got['color'], ref['color']), 'the f16-plane forward was not taken'
    for k in ('depth', 'color', 'loss', 'grad_geo', 'grad_col'):
        assert C.rel_err(got[k].cpu(), ref[k].cpu()) < 3e-5, k
    if 'grad_cam' in got:
        assert C.rel_err(got['grad_cam'].cpu(), ref['grad_cam'].cpu()) < 1e-4
    for k, g in got['grad_params'].items():
        assert C.rel_err(g.cpu(), ref['grad_params'][k].cpu()) < 1e-4, k


@pytest.mark.parametrize('name', ['mapper_color', 'mapper_geometry', 'tracker_color', 'fixed_radius_zero_depth', 'tum_near_pcl',
                                  'expo_mapper_geometry'])
def test_geometry_mma_kernels_agree_with_ffma_kernels(name):
    """Geometry branch with frozen parameters: the warp-level tensor-core kernels (psl_geo_mma.cu, 3xTF32 mma.sync, ReLU-mask save)
    against the fp32 FFMA kernels they replace on that path (PSL_GEO_MMA=0, themselves checked against the oracle): depth, colour,
    loss and every data gradient (features, camera) agree far below the fp32 noise floor of the reference."""
    from point_slam_b200 import ops
    from tests import cases as C
    from tests.gpu_harness import run_case_gpu
    c = C.load_case(name)
    got = run_case_gpu(c, freeze_decoders=True)
    ops.USE_GEO_MMA = False
    try:
        ref = run_case_gpu(c, freeze_decoders=True)
    finally:
        ops.USE_GEO_MMA = True
    assert not torch.equal(got['depth'], ref['depth']), 'the tensor-core geometry forward was not taken'
    keys = ['depth', 'loss', 'grad_geo'] + (['color', 'grad_col'] if c['stage'] == 'color' else [])
    for k in keys:
        e = C.rel_err(got[k].cpu(), ref[k].cpu())
        print(f'{name} {k}: {e:.2e}')
        assert e < 2e-5, k
    if 'grad_cam' in got:
        e = C.rel_err(got['grad_cam'].cpu(), ref['grad_cam'].cpu())
        print(f'{name} grad_cam: {e:.2e}')
        assert e < 1e-4
