"""Forward-only (inference) render timing at BASELINE config C3: 2 M points, 5000 rays x 32 samples (M = 160 k),
tensor-core colour branch on/off, and at the full-image shape (640x480x5 = 1.5 M samples, 500 k points).

    python profiles/prof_forward.py            # prints per-kernel device ms (CUDA events inside the library)
    ncu ... python profiles/prof_forward.py ncu   # a single profiled render between cudaProfilerStart/Stop
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from point_slam_b200 import _lib, ops, synth  # noqa: E402
from tests.test_gpu_scale import _scene  # noqa: E402

DEV = 'cuda:0'
ncu_mode = len(sys.argv) > 1 and sys.argv[1] == 'ncu'


def rays_for(pose, n, S, seed=1):
    depth, color = synth.make_frame(pose)
    _, rq = synth.sobel_radius_map(color)
    o, d = synth.pixel_rays(pose, 480, 640, 517.3, 516.5, 318.6, 255.3)
    pix = np.arange(480 * 640) if n is None else np.random.default_rng(seed).integers(0, 480 * 640, n)
    ro = torch.from_numpy(np.broadcast_to(o, (pix.shape[0], 3)).astype(np.float32).copy()).to(DEV)
    rd = torch.from_numpy(d.reshape(-1, 3)[pix].astype(np.float32)).to(DEV)
    return ro, rd, torch.from_numpy(depth.reshape(-1)[pix]).to(DEV), torch.from_numpy(rq.reshape(-1)[pix]).to(DEV)


def timed(label, npc, dec, ren, ro, rd, gd, dyn, S, reps=5):
    def once():
        with torch.no_grad():
            return ren.render_batch_ray(npc, dec, rd, ro, DEV, 'color', gt_depth=gd, npc_geo_feats=npc.get_geo_feats(),
                                        npc_col_feats=npc.get_col_feats(), cloud_pos=npc.cloud_pos_tensor(), dynamic_r_query=dyn)
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    if ncu_mode:
        torch.cuda.cudart().cudaProfilerStart()
        once()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    _lib.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        once()
    e1.record()
    prof = _lib.timing_collect()
    _lib.timing_enable(False)
    M = ro.shape[0] * S
    wall = e0.elapsed_time(e1) / reps
    ker = {k: v[0] / reps for k, v in prof.items() if v[1]}
    dev_ms = sum(ker.values())
    print(f'{label}: M={M} wall {wall:.3f} ms/render ({M / wall / 1e3:.1f} M samples/s), kernels {dev_ms:.3f} ms '
          f'({M / dev_ms / 1e3:.1f} M samples/s) :: ' + ', '.join(f'{k} {v:.3f}' for k, v in ker.items()))


for tc_on in ((True,) if ncu_mode else (True, False)):
    ops.USE_TENSOR_CORES = tc_on
    cfg, dec, npc, ren, cloud = _scene(2_000_000, S=32)
    ro, rd, gd, dyn = rays_for(synth.trajectory(3)[1], 5000, 32)
    timed(f'C3 2M pts 5000x32 tc={tc_on}', npc, dec, ren, ro, rd, gd, dyn, 32)
    del cfg, dec, npc, ren, cloud
    if not ncu_mode:
        cfg, dec, npc, ren, cloud = _scene(500_000, S=5)
        ro, rd, gd, dyn = rays_for(synth.trajectory(3)[1], None, 5)
        timed(f'full image 640x480x5, 500k pts tc={tc_on}', npc, dec, ren, ro, rd, gd, dyn, 5)
        del cfg, dec, npc, ren, cloud
