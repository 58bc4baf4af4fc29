"""The iteration-shell kernels (csrc/psl_shell.cu) and the fused shells built on them (point_slam_b200/graphed.py) against
their torch restatements: the reference's own op sequences from src/common.py / Tracker.py / Mapper.py as restated in
point_slam_b200/src/common.py and graphed.*_iteration_static (which tests/test_gpu_graphed.py ties to the reference-style
shells).  Tolerances: sampled values exact; gradients / losses 1e-5 relative (fp32 reduction order differs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _lib():
    from point_slam_b200 import _lib as L
    return L, L.load()


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_sample_rays_matches_torch_ops():
    from point_slam_b200.src import common
    from point_slam_b200 import synth
    L, lib = _lib()
    intr = synth.TUM_INTRINSICS
    H, W = intr['H'], intr['W']
    g = torch.Generator(device=DEV).manual_seed(3)
    color = torch.rand(H, W, 3, device=DEV, generator=g)
    depth = torch.rand(H, W, device=DEV, generator=g) * 3
    dyn = torch.rand(H, W, device=DEV, generator=g).double() * 0.1 + 0.02
    cam = torch.tensor([0.9, 0.1, -0.3, 0.2, 0.5, -1.0, 2.0], device=DEV)         # un-normalised quaternion on purpose
    e0, e1, n = 20, 30, 4096
    ww = W - 2 * e1
    pix = torch.randint((H - 2 * e0) * ww, (n,), device=DEV, generator=g)
    ro = torch.empty(n, 3, device=DEV); rd = torch.empty(n, 3, device=DEV); bd = torch.empty(n, device=DEV)
    bc = torch.empty(n, 3, device=DEV); r2 = torch.empty(n, dtype=torch.float64, device=DEV)
    L.check(lib.psl_sample_rays(L.ptr(pix), 1, n, H, W, e0, e1, ww, L.ptr(cam), None, L.ptr(color), L.ptr(depth), L.ptr(dyn),
                                intr['fx'], intr['fy'], intr['cx'], intr['cy'], L.ptr(ro), L.ptr(rd), L.ptr(bd), L.ptr(bc), L.ptr(r2),
                                L.stream()), 'psl_sample_rays')
    jj = torch.div(pix, ww, rounding_mode='floor') + e0
    ii = pix - (jj - e0) * ww + e1
    c2w = common.get_camera_from_tensor(cam)
    ro_t, rd_t = common.get_rays_from_uv(ii.float(), jj.float(), c2w, intr['fx'], intr['fy'], intr['cx'], intr['cy'], DEV)
    assert torch.equal(bd, depth[jj, ii]) and torch.equal(bc, color[jj, ii]) and torch.equal(r2, dyn[jj, ii] ** 2)
    assert torch.equal(ro, ro_t.expand(n, 3))
    assert torch.equal(rd, rd_t)                          # same operation order as ATen's CUDA kernels: bit-identical rays
    # c2w mode, several frames
    K, per = 3, 1000
    c2ws = torch.stack([common.get_camera_from_tensor(cam + 0.1 * k) for k in range(K)])
    colors = torch.rand(K, H, W, 3, device=DEV, generator=g); depths = torch.rand(K, H, W, device=DEV, generator=g)
    pix = torch.randint(H * W, (K, per), device=DEV, generator=g)
    n = K * per
    ro = torch.empty(n, 3, device=DEV); rd = torch.empty(n, 3, device=DEV); bd = torch.empty(n, device=DEV); bc = torch.empty(n, 3, device=DEV)
    L.check(lib.psl_sample_rays(L.ptr(pix), K, per, H, W, 0, 0, W, None, L.ptr(c2ws), L.ptr(colors), L.ptr(depths), None,
                                intr['fx'], intr['fy'], intr['cx'], intr['cy'], L.ptr(ro), L.ptr(rd), L.ptr(bd), L.ptr(bc), None,
                                L.stream()), 'psl_sample_rays')
    jj = torch.div(pix, W, rounding_mode='floor'); ii = pix - jj * W
    kk = torch.arange(K, device=DEV)[:, None].expand(K, per)
    assert torch.equal(bd, depths[kk, jj, ii].reshape(-1)) and torch.equal(bc, colors[kk, jj, ii].reshape(-1, 3))
    for k in range(K):
        o, d = common.get_rays_from_uv(ii[k].float(), jj[k].float(), c2ws[k], intr['fx'], intr['fy'], intr['cx'], intr['cy'], DEV)
        assert torch.equal(ro[k * per:(k + 1) * per], o.expand(per, 3)) and torch.equal(rd[k * per:(k + 1) * per], d)


@pytest.mark.parametrize('n,frac_zero', [(1500, 0.1), (5000, 0.3), (8192, 0.0), (7, 0.5), (64, 1.0)])
def test_depth_gate_matches_masked_stats(n, frac_zero):
    from point_slam_b200 import graphed as G
    L, lib = _lib()
    g = torch.Generator(device=DEV).manual_seed(n)
    d = torch.rand(n, device=DEV, generator=g) * 4 + 0.2
    d[torch.rand(n, device=DEV, generator=g) < 0.02] = 60.0                          # outliers beyond 10 x median
    d[torch.rand(n, device=DEV, generator=g) < frac_zero] = 0.0
    if frac_zero >= 1.0:
        d.zero_()
    depth_in = torch.empty(n, device=DEV); inside = torch.empty(n, dtype=torch.uint8, device=DEV)
    L.check(lib.psl_depth_gate(L.ptr(d), n, L.ptr(depth_in), L.ptr(inside), L.stream()), 'psl_depth_gate')
    valid = d > 0
    ref = valid & (d <= G._masked_stats(d, valid))
    assert torch.equal(inside.bool(), ref)
    assert torch.equal(depth_in, torch.where(ref, d, torch.zeros_like(d)))
    # against the reference's own formulation on the compacted rays (Tracker.py:142-143)
    if bool(valid.any()):
        c = d[valid]
        assert torch.equal(inside.bool()[valid], c <= torch.minimum(10 * c.median(), 1.2 * torch.max(c)))


def _loss_inputs(n, seed, nan_inside):
    g = torch.Generator(device=DEV).manual_seed(seed)
    depth_in = torch.rand(n, device=DEV, generator=g) * 3 + 0.3
    inside = torch.rand(n, device=DEV, generator=g) < 0.85
    depth_in = torch.where(inside, depth_in, torch.zeros_like(depth_in))
    depth = depth_in + 0.05 * torch.randn(n, device=DEV, generator=g)
    depth[5::301] += 30.0                                                            # outliers: beyond 10 x mean
    depth[3::401] += 3000.0                                                          # beyond the 1e3 clamp (tracking)
    var = torch.rand(n, device=DEV, generator=g) * 1e-3 + 1e-6
    nan_at = torch.zeros(n, dtype=torch.bool, device=DEV); nan_at[::97] = True
    if not nan_inside:
        nan_at &= ~inside
    depth[nan_at] = float('nan')
    rgb = torch.rand(n, 3, device=DEV, generator=g)
    b_color = torch.rand(n, 3, device=DEV, generator=g)
    ray_mask = torch.rand(n, device=DEV, generator=g) < 0.9
    return depth_in, inside, ray_mask, depth.requires_grad_(True), var, rgb.requires_grad_(True), b_color


def _torch_loss(G, mode, color, depth_in, inside, ray_mask, depth, var, rgb, b_color, w):
    if mode == 0:                                            # graphed.tracker_iteration_static
        tmp = torch.abs(depth_in - depth) / torch.sqrt(var + 1e-10)
        with torch.no_grad():
            ok = inside & (~torch.isnan(depth)) & (~torch.isnan(var))
            mean_tmp = G._msum(tmp, inside) / inside.sum().clamp_min(1)
            mask = ok & (tmp < 10 * mean_tmp)
        return G._msum(torch.clamp(tmp, min=0.0, max=1e3), mask) + w * G._msum(torch.abs(b_color - rgb), mask[:, None].expand(-1, 3))
    m = inside & ray_mask & (~torch.isnan(depth))            # graphed.mapper_iteration_static
    loss = G._msum(torch.abs(depth_in - depth), m)
    if color:
        loss = loss + w * G._msum(torch.abs(b_color - rgb), m[:, None].expand(-1, 3))
    return loss


@pytest.mark.parametrize('mode,color,nan_inside', [(0, True, False), (1, True, True), (1, False, True), (0, True, True)])
def test_shell_loss_and_gradients_match_autograd(mode, color, nan_inside):
    """nan_inside with the tracking loss: a NaN depth among the gated rays poisons the mean exactly as in the reference
    (tmp.mean() over the batch, Tracker.py:171) -> every ray is masked, loss 0."""
    from point_slam_b200 import graphed as G
    n = 3000
    w = 0.5 if mode == 0 else 0.1
    depth_in, inside, ray_mask, depth, var, rgb, b_color = _loss_inputs(n, 11 + mode, nan_inside)
    loss = _torch_loss(G, mode, color, depth_in, inside, ray_mask, depth, var, rgb, b_color, w)
    loss.backward()
    if mode == 0 and nan_inside:
        assert float(loss) == 0.0
    else:
        assert float(loss) > 0.0
    _check_loss(mode, color, n, depth_in, inside, ray_mask, depth, var, rgb, b_color, w, loss)


def _check_loss(mode, color, n, depth_in, inside, ray_mask, depth, var, rgb, b_color, w, loss):
    L, lib = _lib()
    out = torch.zeros((), device=DEV); dd = torch.empty(n, device=DEV); dc = torch.empty(n, 3, device=DEV) if color else None
    ins = inside.to(torch.uint8).contiguous(); rm = ray_mask.to(torch.uint8).contiguous()
    L.check(lib.psl_shell_loss(mode, n, L.ptr(depth_in), L.ptr(ins), L.ptr(rm), L.ptr(depth.detach()), L.ptr(var), L.ptr(rgb.detach()),
                               L.ptr(b_color), w, L.ptr(out), L.ptr(dd), L.ptr(dc), L.stream()), 'psl_shell_loss')
    assert abs(float(out) - float(loss)) <= 1e-5 * abs(float(loss)), (float(out), float(loss))
    gd = torch.nan_to_num(depth.grad, nan=0.0)               # autograd leaves NaN where the masked-out input was NaN
    assert float((dd - gd).abs().max()) <= 1e-6 * float(gd.abs().max().clamp_min(1e-30))
    if color:
        assert torch.equal(dc, rgb.grad if rgb.grad is not None else torch.zeros_like(dc))


@pytest.mark.parametrize('mode,color', [(0, True), (1, True), (1, False)])
def test_render_tail_equals_the_four_kernels_it_fuses(mode, color):
    """psl_render_tail == psl_composite_fwd + psl_ray_mask + psl_shell_loss + psl_composite_bwd (same per-ray arithmetic, same
    summation order): outputs, loss and d_raw, incl. rays without neighbours, NaN depth inputs and rays outside the gate."""
    from point_slam_b200 import _lib as L
    lib = L.load()
    g = torch.Generator(device=DEV).manual_seed(5 + mode)
    n, S = 4321, 5
    raw = torch.randn(n * S, 4, device=DEV, generator=g)
    raw[:, 3] *= 30
    has_nb = (torch.rand(n * S, device=DEV, generator=g) > 0.2).to(torch.uint8)
    z = (torch.rand(n, S, device=DEV, generator=g) * 0.1).cumsum(1) + 1.0
    depth_in = 1.0 + 0.3 * torch.rand(n, device=DEV, generator=g)
    inside = (torch.rand(n, device=DEV, generator=g) > 0.1).to(torch.uint8)
    b_color = torch.rand(n, 3, device=DEV, generator=g)
    f = lambda *shape, dt=torch.float32: torch.empty(*shape, dtype=dt, device=DEV)
    depth, var, rgb, mask, loss, d_raw = f(n), f(n), f(n, 3), f(n, dt=torch.uint8), f(()), f(n * S, 4)
    L.check(lib.psl_composite_fwd(L.ptr(raw), L.ptr(has_nb), L.ptr(z), n, S, 0.1, L.ptr(depth), L.ptr(var), L.ptr(rgb), None, L.stream()), 'fwd')
    L.check(lib.psl_ray_mask(L.ptr(has_nb), n, S, int(S / 2 + 1), L.ptr(mask), L.stream()), 'mask')
    d_depth, d_rgb = f(n), (f(n, 3) if color else None)
    L.check(lib.psl_shell_loss(mode, n, L.ptr(depth_in), L.ptr(inside), L.ptr(mask), L.ptr(depth), L.ptr(var), L.ptr(rgb), L.ptr(b_color),
                               0.3, L.ptr(loss), L.ptr(d_depth), L.ptr(d_rgb), L.stream()), 'loss')
    L.check(lib.psl_composite_bwd(L.ptr(raw), L.ptr(has_nb), L.ptr(z), n, S, 0.1, L.ptr(d_depth), None, L.ptr(d_rgb), L.ptr(d_raw), L.stream()), 'bwd')
    depth2, var2, rgb2, mask2, loss2, d_raw2 = f(n), f(n), f(n, 3), f(n, dt=torch.uint8), f(()), f(n * S, 4)
    ws = torch.zeros(lib.psl_render_tail_ws_bytes(n), dtype=torch.uint8, device=DEV)
    for _ in range(2):                                   # second launch: the ticket was re-armed by the first
        L.check(lib.psl_render_tail(mode, n, S, 0.1, int(S / 2 + 1), L.ptr(raw), L.ptr(has_nb), L.ptr(z), L.ptr(depth_in), L.ptr(inside),
                                L.ptr(b_color) if color else None, 0.3, L.ptr(depth2), L.ptr(var2), L.ptr(rgb2), L.ptr(mask2), L.ptr(loss2),
                                L.ptr(d_raw2), L.ptr(ws), ws.numel(), L.stream()), 'tail')
    torch.cuda.synchronize()
    assert torch.equal(depth, depth2) and torch.equal(var, var2) and torch.equal(rgb, rgb2) and torch.equal(mask, mask2)
    assert float((loss - loss2).abs()) <= 1e-6 * float(loss.abs()), (float(loss), float(loss2))
    assert float(d_raw.abs().max()) > 0
    assert float((d_raw - d_raw2).abs().max()) <= 1e-6 * float(d_raw.abs().max())


def test_pose_bwd_matches_autograd():
    from point_slam_b200.src import common
    from point_slam_b200 import synth
    L, lib = _lib()
    intr = synth.TUM_INTRINSICS
    H, W = intr['H'], intr['W']
    g = torch.Generator(device=DEV).manual_seed(8)
    cam = torch.tensor([0.8, -0.2, 0.4, 0.1, 0.3, 0.2, -0.7], device=DEV, requires_grad=True)
    e0, e1, n = 100, 100, 1500
    ww = W - 2 * e1
    pix = torch.randint((H - 2 * e0) * ww, (n,), device=DEV, generator=g)
    jj = torch.div(pix, ww, rounding_mode='floor') + e0
    ii = pix - (jj - e0) * ww + e1
    ro, rd = common.get_rays_from_uv(ii.float(), jj.float(), common.get_camera_from_tensor(cam), intr['fx'], intr['fy'], intr['cx'],
                                     intr['cy'], DEV)
    g_o = torch.randn(n, 3, device=DEV, generator=g); g_d = torch.randn(n, 3, device=DEV, generator=g)
    ((ro.expand(n, 3) * g_o).sum() + (rd * g_d).sum()).backward()
    d_cam = torch.empty(7, device=DEV)
    L.check(lib.psl_pose_bwd(L.ptr(pix), n, e0, e1, ww, intr['fx'], intr['fy'], intr['cx'], intr['cy'], L.ptr(cam.detach()), L.ptr(g_o),
                             L.ptr(g_d), L.ptr(d_cam), L.stream()), 'psl_pose_bwd')
    assert _rel(d_cam, cam.grad) < 1e-5, (d_cam, cam.grad)


def test_adam_rows_matches_torch_adam():
    from point_slam_b200 import graphed as G
    g = torch.Generator(device=DEV).manual_seed(2)
    N, U = 5000, 1200
    full = torch.randn(N, 32, device=DEV, generator=g)
    idx = torch.randperm(N, device=DEV, generator=g)[:U]
    rows = torch.full((2048,), -1, dtype=torch.int64, device=DEV); rows[:U] = idx
    ref = full[idx].clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.005)
    ad = G.AdamRows(2048, 32, DEV, 0.005)
    mine = full.clone()
    for it in range(25):
        gr = torch.randn(U, 32, device=DEV, generator=g) * (10.0 ** float(it % 5 - 3))
        gr[::7] = 0.0
        ref.grad = gr.clone()
        opt.step()
        ad.grad[:U] = gr
        ad.step(mine, rows)
        assert float(ad.grad.abs().max()) == 0.0             # consumed gradients are cleared
    assert _rel(mine[idx], ref.detach()) < 2e-6
    untouched = torch.ones(N, dtype=torch.bool, device=DEV); untouched[idx] = False
    assert torch.equal(mine[untouched], full[untouched])
    # dense single row (the pose)
    cam = torch.randn(7, device=DEV, generator=g); ref = cam.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.002); ad = G.AdamRows(1, 7, DEV, 0.002)
    for it in range(40):
        gr = torch.randn(7, device=DEV, generator=g)
        ref.grad = gr.clone(); opt.step()
        ad.grad.copy_(gr.view(1, 7)); ad.step(cam, None, zero_grad=False)
    assert _rel(cam, ref.detach()) < 2e-6


def _scene():
    import bench
    return bench, bench.GpuScene(0, DEV, bench.CONFIGS['c2'], 200000, 1)


def test_fused_tracker_iteration_matches_static_shell():
    from point_slam_b200 import graphed as G, ops
    bench, scene = _scene()
    cur = scene.resident[0]
    npc, dec, ren = scene.npc, scene.decoders, scene.renderer
    cam0 = bench.cam_tensor_from_c2w(scene.frames_host[bench.N_KEYFRAMES]['c2w'], 0.01, scene.rng).to(DEV)
    cam = cam0.clone().requires_grad_(True)
    torch.manual_seed(5)
    loss_s = G.tracker_iteration_static(ren, npc, dec, cam, cur['color'], cur['depth'], cur['dyn_r_query'], bench.INTR, 1500, DEV,
                                        npc.get_geo_feats(), npc.get_col_feats(), npc.cloud_pos_tensor(), (100, 100))
    torch.manual_seed(5)
    d_cam = torch.zeros(1, 7, device=DEV); loss_f = torch.zeros((), device=DEV)
    pack = ops.PackedDecoder(DEV).pack(dec.kernel_params())
    G.tracker_iteration_fused(ren, npc, dec, cam0.clone(), d_cam, cur['color'], cur['depth'], cur['dyn_r_query'], bench.INTR, 1500, DEV,
                              npc.get_geo_feats(), npc.get_col_feats(), npc.cloud_pos_tensor(), (100, 100), loss_f, pack=pack,
                              prepacked=True)
    assert abs(float(loss_f) - float(loss_s)) / abs(float(loss_s)) < 1e-5, (float(loss_f), float(loss_s))
    assert _rel(d_cam.view(-1), cam.grad) < 1e-4, (d_cam, cam.grad)


@pytest.mark.parametrize('stage', ['geometry', 'color'])
def test_fused_mapper_iteration_matches_static_shell(stage):
    from point_slam_b200 import graphed as G, iteration as IT
    bench, scene = _scene()
    cur = scene.resident[0]
    npc, dec, ren = scene.npc, scene.decoders, scene.renderer
    idx = IT.frustum_indices(npc.cloud_pos_tensor(), cur['c2w'], bench.INTR)
    kfl = [cur] + scene.keyframes
    # torch-shell reference: slices + index_put + autograd
    gm = G.GraphedMapper(ren, npc, dec, bench.INTR, 5000, DEV)
    gm.begin_frame(idx, kfl)
    st = gm.state
    st.optimizer.zero_grad(set_to_none=True)
    torch.manual_seed(9)
    loss_s = G.mapper_iteration_static(ren, npc, dec, st, gm.keyframes, bench.INTR, 5000, DEV, stage, npc.cloud_pos_tensor())
    U = idx.shape[0]
    g_geo = st.geo.grad[:U].clone()
    g_col = st.col.grad[:U].clone() if stage == 'color' else None
    g_dec = {n: p.grad.clone() for n, p in dec.color_decoder.named_parameters() if p.grad is not None}
    for p in dec.parameters():
        p.grad = None
    # fused
    fm = G.FusedMapper(ren, npc, dec, bench.INTR, 5000, DEV)
    fm.begin_frame(idx, kfl)
    torch.manual_seed(9)
    G.mapper_iteration_fused(ren, npc, dec, fm, fm.keyframes, bench.INTR, 5000, DEV, stage, npc.cloud_pos_tensor(), fm.loss,
                             apply_adam=False)
    assert abs(float(fm.loss) - float(loss_s)) / abs(float(loss_s)) < 1e-5, (float(fm.loss), float(loss_s))
    assert _rel(fm.adam_geo.grad[:U], g_geo) < 1e-5
    assert float(fm.adam_geo.grad[U:].abs().max()) == 0.0
    if stage == 'color':
        assert _rel(fm.adam_col.grad[:U], g_col) < 1e-5
        assert len(g_dec) > 0
        for n, p in dec.color_decoder.named_parameters():
            if n in g_dec:
                assert _rel(p.grad, g_dec[n]) < 1e-4, n


def test_fused_graphs_run_and_optimise():
    from point_slam_b200 import graphed as G, iteration as IT
    bench, scene = _scene()
    cur = scene.resident[0]
    npc, dec, ren = scene.npc, scene.decoders, scene.renderer
    ft = G.FusedTracker(ren, npc, dec, bench.INTR, 1500, DEV, edge=(100, 100), separate_lr=False)     # one lr, like the torch-shell Adam below
    cam0 = bench.cam_tensor_from_c2w(scene.frames_host[bench.N_KEYFRAMES]['c2w'], 0.02, scene.rng).to(DEV)
    ft.load_frame(cur['color'], cur['depth'], cur['dyn_r_query'], cam0)
    l0 = float(ft.run(1))
    l1 = float(ft.run(40))
    assert np.isfinite(l0) and np.isfinite(l1), (l0, l1)      # (random-init decoder: the loss value itself says little)
    assert int(ft.adam.t) == 41
    # same trajectory as the torch-shell graphs (same RNG stream, same Adam).  The two shells are bit-identical until a
    # rounding difference of the pose-gradient reduction flips one ulp of the pose (about once per 500 component-steps,
    # measured); the objective is discontinuous in the pose (radius cut-off of the kNN weights, outlier masks), so after
    # such a flip the trajectories separate within ~10 iterations.  Hence: tight agreement over a short horizon, same
    # ballpark over the full 41 iterations.
    gt = G.GraphedTracker(ren, npc, dec, bench.INTR, 1500, DEV, edge=(100, 100))
    gt.capture()
    for n_it, tol in ((5, 0.05), (41, None)):
        poses = []
        for tr in (ft, gt):
            tr.load_frame(cur['color'], cur['depth'], cur['dyn_r_query'], cam0)
            torch.manual_seed(77)
            tr.run(n_it)
            poses.append(tr.cam.detach().clone())
        disp = [float((p - cam0).abs().max()) for p in poses]
        assert min(disp) > 1e-3
        if tol is not None:
            assert float((poses[0] - poses[1]).abs().max()) < tol * disp[1] + 1e-6, poses
        else:
            assert 0.3 < disp[0] / disp[1] < 3.0 and all(bool(torch.isfinite(p).all()) for p in poses), poses
    fm = G.FusedMapper(ren, npc, dec, bench.INTR, 5000, DEV)
    idx = IT.frustum_indices(npc.cloud_pos_tensor(), cur['c2w'], bench.INTR)
    fm.begin_frame(idx, [cur] + scene.keyframes)
    g0 = fm.npc_geo.clone(); c0 = fm.npc_col.clone()
    w0 = [p.detach().clone() for p in dec.color_decoder.parameters()]
    la = float(fm.run('geometry', 3))
    lb = float(fm.run('geometry', 20))
    assert torch.equal(fm.npc_col, c0)                       # colour features do not move in the geometry stage
    lc = float(fm.run('color', 2))
    ld = float(fm.run('color', 20))
    assert all(np.isfinite(x) for x in (la, lb, lc, ld)) and lb < la and ld < lc, (la, lb, lc, ld)
    assert int(fm.adam_geo.t) == 45 and int(fm.adam_col.t) == 22
    moved = (fm.npc_geo - g0).abs().sum(1) > 0
    sel = torch.zeros_like(moved); sel[idx] = True
    assert bool(moved.any()) and not bool((moved & ~sel).any())          # only frustum rows are optimised
    assert any(not torch.equal(a, p.detach()) for a, p in zip(w0, dec.color_decoder.parameters()))
    # the rows are optimised in place in the cloud's own tensors: write_back (Mapper.py:605-610) has nothing left to copy
    assert fm.npc_geo.data_ptr() == npc.get_geo_feats().data_ptr() and torch.equal(fm.write_back(), idx)
    assert not torch.equal(g0, npc.get_geo_feats())
    n_graphs = len(fm.graphs)
    fm.begin_frame(idx[: idx.shape[0] // 2], [cur] + scene.keyframes)    # a different frustum re-uses the graphs
    le = float(fm.run('color', 3))
    assert np.isfinite(le) and len(fm.graphs) == n_graphs


def test_feat_scatter_mapped_equals_dense_scatter_on_selected_rows():
    """psl_feat_scatter_mapped writes the compact gradient of a row subset; per row the pairs are summed in the same (pair)
    order as the dense scatter, so the selected rows are bit-identical and unselected points are dropped."""
    L, lib = _lib()
    g = torch.Generator(device=DEV).manual_seed(21)
    N, M, U = 20000, 6000, 3000
    I = torch.randint(0, N, (M, 8), device=DEV, generator=g, dtype=torch.int32)
    I[torch.rand(M, 8, device=DEV, generator=g) < 0.2] = -1
    wn = torch.rand(M, 8, device=DEV, generator=g)
    wn[I < 0] = 0.0
    wn[torch.rand(M, 8, device=DEV, generator=g) < 0.1] = 0.0
    d_cg = torch.randn(M, 32, device=DEV, generator=g)
    d_colpair = torch.randn(M, 8, 32, device=DEV, generator=g)
    ws_bytes = lib.psl_feat_scatter_ws_bytes(M)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    dg = torch.zeros(N, 32, device=DEV); dc = torch.zeros(N, 32, device=DEV)
    L.check(lib.psl_feat_scatter(L.ptr(I), M, N, L.ptr(wn), L.ptr(d_cg), L.ptr(d_colpair), None, L.ptr(dg), L.ptr(dc), L.ptr(ws),
                                 ws_bytes, L.stream()), 'psl_feat_scatter')
    rows = torch.randperm(N, device=DEV, generator=g)[:U]
    row_map = torch.full((N,), -1, dtype=torch.int32, device=DEV)
    row_map[rows] = torch.arange(U, dtype=torch.int32, device=DEV)
    cap = 4096                                               # capacity larger than U: the tail stays zero
    mg = torch.zeros(cap, 32, device=DEV); mc = torch.zeros(cap, 32, device=DEV)
    L.check(lib.psl_feat_scatter_mapped(L.ptr(I), M, L.ptr(row_map), cap, L.ptr(wn), L.ptr(d_cg), L.ptr(d_colpair), None, L.ptr(mg),
                                        L.ptr(mc), L.ptr(ws), ws_bytes, L.stream()), 'psl_feat_scatter_mapped')
    assert torch.equal(mg[:U], dg[rows]) and torch.equal(mc[:U], dc[rows])
    assert float(mg[U:].abs().max()) == 0.0 and float(mc[U:].abs().max()) == 0.0
    assert float(dg.abs().sum()) > 0
    # reference semantics: index_put_(accumulate) of w * d_cg into feats[I] (decoder.py:164), fp64 for an order-free check
    ref = torch.zeros(N, 32, device=DEV, dtype=torch.float64)
    valid = (I >= 0) & (wn != 0)
    contrib = (wn[..., None].double() * d_cg[:, None, :].double())[valid]
    ref.index_put_((I[valid].long(),), contrib, accumulate=True)
    assert float((dg.double() - ref).abs().max()) < 1e-5 * float(ref.abs().max())


def test_feat_scatter_long_segments_sum_in_pair_order():
    """Few points, many pairs per point (segments of several hundred pairs: multi-chunk path of k_scatter_segments): every row is
    the fused-multiply-add chain over its pairs in ascending pair order -- checked bit for bit against a float64-emulated fma
    chain on the host -- for the per-pair colour gradient (rel-pos on), the per-sample one (rel-pos off) and the geometry one."""
    L, lib = _lib()
    g = torch.Generator(device=DEV).manual_seed(33)
    N, M = 37, 3000
    I = torch.randint(0, N, (M, 8), device=DEV, generator=g, dtype=torch.int32)
    I[torch.rand(M, 8, device=DEV, generator=g) < 0.1] = -1
    I[:, 0] = 5                                              # one very long segment (3000 pairs)
    wn = torch.rand(M, 8, device=DEV, generator=g)
    wn[I < 0] = 0.0
    wn[torch.rand(M, 8, device=DEV, generator=g) < 0.05] = 0.0
    d_cg = torch.randn(M, 32, device=DEV, generator=g)
    d_cc = torch.randn(M, 32, device=DEV, generator=g)
    d_colpair = torch.randn(M, 8, 32, device=DEV, generator=g)
    ws_bytes = lib.psl_feat_scatter_ws_bytes(M)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    out = {}
    for name, pair, cc in (('rel', d_colpair, None), ('norel', None, d_cc)):
        dg = torch.zeros(N, 32, device=DEV); dc = torch.zeros(N, 32, device=DEV)
        L.check(lib.psl_feat_scatter(L.ptr(I), M, N, L.ptr(wn), L.ptr(d_cg), L.ptr(pair), L.ptr(cc), L.ptr(dg), L.ptr(dc), L.ptr(ws),
                                     ws_bytes, L.stream()), 'psl_feat_scatter')
        out[name] = (dg.cpu().numpy(), dc.cpu().numpy())
    assert np.array_equal(out['rel'][0], out['norel'][0])
    Ih, wh = I.cpu().numpy().reshape(-1), wn.cpu().numpy().reshape(-1)
    cg, cc_h, cp = d_cg.cpu().numpy(), d_cc.cpu().numpy(), d_colpair.cpu().numpy().reshape(-1, 32)
    fma = lambda a, b, c: (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)
    for row in (5, 0, 11, 36):
        ag = np.zeros(32, np.float32); ac_rel = np.zeros(32, np.float32); ac_no = np.zeros(32, np.float32)
        pairs = np.nonzero((Ih == row) & (wh != 0))[0]
        assert pairs.size > 64
        for pid in pairs:                                    # ascending pair id == the stable sort's order inside a segment
            w = np.full(32, wh[pid], np.float32)
            ag = fma(w, cg[pid >> 3], ag)
            ac_rel = (ac_rel + cp[pid]).astype(np.float32)
            ac_no = fma(w, cc_h[pid >> 3], ac_no)
        assert np.array_equal(out['rel'][0][row], ag), row
        assert np.array_equal(out['rel'][1][row], ac_rel), row
        assert np.array_equal(out['norel'][1][row], ac_no), row


def test_shell_entry_points_reject_bad_arguments():
    L, lib = _lib()
    d = torch.zeros(9000, device=DEV); o = torch.zeros(9000, device=DEV); m = torch.zeros(9000, dtype=torch.uint8, device=DEV)
    assert lib.psl_depth_gate(L.ptr(d), 9000, L.ptr(o), L.ptr(m), L.stream()) != 0
    assert b'8192' in lib.psl_last_error()
    assert lib.psl_depth_gate(None, 10, L.ptr(o), L.ptr(m), L.stream()) != 0
    loss = torch.zeros((), device=DEV)
    assert lib.psl_shell_loss(0, 10, L.ptr(d), L.ptr(m), None, L.ptr(d), None, None, None, 0.5, L.ptr(loss), L.ptr(o), None, L.stream()) != 0   # tracking needs var
    assert lib.psl_shell_loss(1, 10, L.ptr(d), L.ptr(m), None, L.ptr(d), None, None, None, 0.1, L.ptr(loss), L.ptr(o), None, L.stream()) != 0   # mapping needs ray_mask
    pix = torch.zeros(4, dtype=torch.int64, device=DEV)
    assert lib.psl_sample_rays(L.ptr(pix), 1, 4, 480, 640, 0, 0, 640, None, None, L.ptr(d), L.ptr(d), None, 1., 1., 0., 0., L.ptr(o), L.ptr(o),
                               L.ptr(o), L.ptr(o), None, L.stream()) != 0                                                     # neither cam nor c2w
    torch.cuda.synchronize()


def test_iteration_graphs_survive_cloud_growth():
    """add_neural_points between frames (Mapper.py:317,328) must not force a re-capture: the hash is rebuilt in place in its
    capacity buffers and the kernels read its size from device memory (ops.SpatialHash / psl_grid_meta).  The replayed graph
    must see the NEW cloud: same result as a tracker captured from scratch on the grown cloud."""
    from point_slam_b200 import graphed as G
    from point_slam_b200.src import common
    bench, scene = _scene()
    cur = scene.resident[0]
    npc, dec, ren = scene.npc, scene.decoders, scene.renderer
    ft = G.FusedTracker(ren, npc, dec, bench.INTR, 1500, DEV, edge=(100, 100))
    cam0 = bench.cam_tensor_from_c2w(scene.frames_host[bench.N_KEYFRAMES]['c2w'], 0.02, scene.rng).to(DEV)
    ft.load_frame(cur['color'], cur['depth'], cur['dyn_r_query'], cam0)
    ft.run(3)
    assert ft.captures == 1
    # carve a hole into the view by adding points in front of the camera: a batch of rays at 60 % of the sensor depth
    n0 = npc.pts_num()
    pix = torch.randint(0, bench.INTR['H'] * bench.INTR['W'], (4000,), device=DEV)
    j, i = pix // bench.INTR['W'], pix % bench.INTR['W']
    ro, rd = common.get_rays_from_uv(i.float(), j.float(), cur['c2w'], bench.INTR['fx'], bench.INTR['fy'], bench.INTR['cx'], bench.INTR['cy'], DEV)
    k = npc.add_neural_points(ro, rd, cur['depth'][j, i] * 0.6, cur['color'][j, i])
    assert int(k) > 1000 and npc.pts_num() == n0 + 3 * int(k)
    gen = npc.spatial_hash().build_gen
    ft.load_frame(cur['color'], cur['depth'] * 0.6, cur['dyn_r_query'], cam0)      # the frame now looks at the new points
    torch.manual_seed(3)
    ft.run(5)
    assert ft.captures == 1, 'the tracker graph was re-captured although no buffer was re-allocated'
    fresh = G.FusedTracker(ren, npc, dec, bench.INTR, 1500, DEV, edge=(100, 100))
    fresh.load_frame(cur['color'], cur['depth'] * 0.6, cur['dyn_r_query'], cam0)
    fresh.capture()
    torch.manual_seed(3)
    fresh.run(5)
    assert npc.spatial_hash().build_gen == gen
    assert torch.equal(ft.cam, fresh.cam) and torch.equal(ft.loss, fresh.loss), (ft.cam, fresh.cam)
    # a frame that adds nothing leaves the hash alone (no rebuild)
    k2 = npc.add_neural_points(ro, rd, cur['depth'][j, i] * 0.6, cur['color'][j, i])
    assert int(k2) == 0 and npc.spatial_hash().build_gen == gen


def test_single_process_scheduler_runs_a_short_sequence():
    """PointSLAM (point_slam_b200/slam.py): the reference's tracker / mapper hand-shake as one loop on one GPU.  Empty cloud,
    eight frames of a smooth synthetic trajectory, reduced iteration counts: frame 0 creates the map, later frames are tracked and
    every second one is mapped; the cloud grows, the iteration graphs are captured once per stage, the poses stay close to the
    ground truth."""
    import types
    from point_slam_b200 import synth
    from point_slam_b200.default_config import make_cfg
    from point_slam_b200.slam import PointSLAM
    from point_slam_b200.src.conv_onet import config as model_config
    from point_slam_b200.src.neural_point import NeuralPointCloud
    from point_slam_b200.src.utils.Renderer import Renderer
    import bench
    intr = bench.INTR
    cfg = make_cfg('replica', DEV, **{'mapping.every_frame': 2, 'mapping.keyframe_every': 2, 'mapping.mapping_window_size': 4,
                                      'mapping.pixels': 2000, 'mapping.pixels_adding': 6000, 'mapping.pixels_based_on_color_grad': 500,
                                      'tracking.pixels': 1500, 'tracking.iters': 40, 'tracking.ignore_edge_W': 40, 'tracking.ignore_edge_H': 40})
    dec = bench.build_decoders(cfg, DEV)
    npc = NeuralPointCloud(cfg)
    npc.reserve(200000)
    ren = Renderer(cfg, None, types.SimpleNamespace(**{k: intr[k] for k in ('H', 'W', 'fx', 'fy', 'cx', 'cy')}))
    ren.sigmoid_coefficient = 0.1
    slam = PointSLAM(cfg, ren, npc, dec, intr, DEV, mapping_iters=100, first_iters=500)
    poses = synth.trajectory(400, seed=3)[:8]              # 8 consecutive frames of a 400-frame loop: ~1 cm / frame
    errs = []
    for idx in range(8):
        depth, color = synth.make_frame(poses[idx], intr)
        gt = torch.from_numpy(poses[idx]).float().to(DEV)
        est = slam.process_frame(idx, torch.from_numpy(color).to(DEV), torch.from_numpy(depth).to(DEV), gt)
        errs.append(float((est[:3, 3] - gt[:3, 3]).norm()))
    torch.cuda.synchronize()
    log = slam.log
    assert [e['mapped'] for e in log] == [True, False, True, False, True, False, True, False]
    assert log[0]['added'] > 3000 and npc.pts_num() == 3 * sum(e.get('added', 0) for e in log) > 9000
    assert npc.index_ntotal() == npc.pts_num()
    print('pose errors (m):', [round(e, 4) for e in errs])
    assert all(np.isfinite(errs)) and max(errs) < 0.03, errs               # tracked frames stay within 3 cm of the ground truth (measured: 1.2 cm)
    assert slam.tracker.captures == 1                                       # the cloud grew under the captured graphs
    assert slam.mapper.captures <= 4                                        # init / stage learning rates x two stages
