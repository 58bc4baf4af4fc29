"""Static-shape / CUDA-graph iteration shells (point_slam_b200/graphed.py) against the reference-style shells."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _setup():
    import bench
    scene = bench.GpuScene(0, DEV, bench.CONFIGS['c2'], 200000, 1)
    return bench, scene


def test_static_tracker_iteration_matches_reference_style_shell():
    from point_slam_b200 import iteration as IT, graphed as G
    bench, scene = _setup()
    cur = scene.resident[0]
    npc, dec, ren = scene.npc, scene.decoders, scene.renderer
    cam0 = bench.cam_tensor_from_c2w(scene.frames_host[bench.N_KEYFRAMES]['c2w'], 0.01, scene.rng).to(DEV)
    out = {}
    for kind in ('ref', 'static'):
        cam = cam0.clone().requires_grad_(True)
        opt = torch.optim.SGD([cam], lr=0.0)                 # keep the pose fixed: compare loss and gradient only
        torch.manual_seed(5)
        if kind == 'ref':
            loss, _ = IT.tracker_iteration(ren.render_batch_ray, npc, dec, cam, opt, cur['color'], cur['depth'], cur['dyn_r_query'],
                                           bench.INTR, 1500, DEV, npc.get_geo_feats(), npc.get_col_feats(), npc.cloud_pos_tensor(),
                                           edge=(100, 100))
            out[kind] = (float(loss), None)
        else:
            loss = G.tracker_iteration_static(ren, npc, dec, cam, cur['color'], cur['depth'], cur['dyn_r_query'], bench.INTR, 1500,
                                              DEV, npc.get_geo_feats(), npc.get_col_feats(), npc.cloud_pos_tensor(), (100, 100))
            out[kind] = (float(loss), cam.grad.clone())
    # the reference-style shell zeroes the gradient after its optimizer step; recompute it for the comparison
    cam = cam0.clone().requires_grad_(True)
    torch.manual_seed(5)

    class Keep(torch.optim.SGD):
        def zero_grad(self, *a, **k):
            pass
    IT.tracker_iteration(ren.render_batch_ray, npc, dec, cam, Keep([cam], lr=0.0), cur['color'], cur['depth'], cur['dyn_r_query'],
                         bench.INTR, 1500, DEV, npc.get_geo_feats(), npc.get_col_feats(), npc.cloud_pos_tensor(), edge=(100, 100))
    g_ref = cam.grad
    assert abs(out['ref'][0] - out['static'][0]) / abs(out['ref'][0]) < 1e-5
    assert float((g_ref - out['static'][1]).abs().max() / g_ref.abs().max()) < 1e-4


def test_graphed_tracker_and_mapper_run_and_optimise():
    from point_slam_b200 import iteration as IT, graphed as G
    bench, scene = _setup()
    cur = scene.resident[0]
    npc, dec, ren = scene.npc, scene.decoders, scene.renderer
    gt = G.GraphedTracker(ren, npc, dec, bench.INTR, 1500, DEV, edge=(100, 100))
    cam0 = bench.cam_tensor_from_c2w(scene.frames_host[bench.N_KEYFRAMES]['c2w'], 0.02, scene.rng).to(DEV)
    gt.load_frame(cur['color'], cur['depth'], cur['dyn_r_query'], cam0)
    # The synthetic cloud carries RANDOM features (synth.make_features), so the render does not depict the frame and the tracking
    # loss has no minimum near the true pose: "the loss goes down" is not a property of this setting (measured: on one fixed
    # pixel batch it moves by +-3 % over 41 iterations).  What must hold: finite losses, Adam-bounded steps (|step| <= ~lr per
    # iteration), and the exact loss / gradient agreement with the reference-style shell checked in the test above.
    l0 = float(gt.run(1))
    l1 = float(gt.run(40))
    torch.cuda.synchronize()
    assert np.isfinite(l0) and np.isfinite(l1) and 0.5 * l0 < l1 < 2.0 * l0, (l0, l1)
    assert float((gt.cam.detach() - cam0).abs().max()) <= 41 * 0.002 * 1.5
    assert float((gt.cam.detach() - cam0).abs().max()) > 0
    # mapper
    gm = G.GraphedMapper(ren, npc, dec, bench.INTR, 5000, DEV)
    idx = IT.frustum_indices(npc.cloud_pos_tensor(), cur['c2w'], bench.INTR)
    gm.begin_frame(idx, [cur] + scene.keyframes)
    state = gm.state
    g0 = state.geo.detach().clone()
    la = float(gm.run('geometry', 3))
    lb = float(gm.run('geometry', 20))
    lc = float(gm.run('color', 10))
    torch.cuda.synchronize()
    assert np.isfinite(la) and np.isfinite(lb) and np.isfinite(lc) and lb < la
    assert float((state.geo.detach() - g0).abs().max()) > 0
    assert float(state.geo.detach()[gm.n_used:].abs().max()) == 0.0          # padded slots never move
    before = npc.get_geo_feats().clone()
    gm.write_back()
    assert not torch.equal(before, npc.get_geo_feats())
    gm.begin_frame(idx[: idx.shape[0] // 2], [cur] + scene.keyframes)         # a different frustum: same graphs are reused
    n_graphs = len(gm.graphs)
    ld = float(gm.run('color', 5))
    assert np.isfinite(ld) and len(gm.graphs) == n_graphs
