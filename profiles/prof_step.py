"""Profiling driver: the bench scene, a few tracking / mapping iterations (the static-shape shells that bench.py replays as
CUDA graphs, launched eagerly here so that ncu sees every kernel) inside a cudaProfilerStart/Stop window.

    ncu --profile-from-start off ... python profiles/prof_step.py [n_track n_map]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from point_slam_b200 import iteration as IT  # noqa: E402

n_track = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_map = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = 'cuda:0'
scene = bench.GpuScene(0, dev, bench.CONFIGS[os.environ.get('PSL_PROF_CONFIG', 'c2')], 500000, 1)
fh = scene.frames_host[bench.N_KEYFRAMES]
src = scene.resident[0]
tr, mp = scene.tracker, scene.mapper
tr.load_frame(src['color'], src['depth'], src['dyn_r_query'], bench.cam_tensor_from_c2w(fh['c2w'], 0.01, scene.rng).to(dev))
cur = dict(color=tr.color, depth=tr.depth, dyn_r_query=tr.dyn, c2w=src['c2w'])
INTR = bench.INTR


def select():
    return scene.ops.frustum_select(scene.npc.cloud_pos_tensor(), fh['c2w'], tr.depth, INTR['H'], INTR['W'], INTR['fx'], INTR['fy'],
                                    INTR['cx'], INTR['cy'], edge=-4)


idx = select()
mp.begin_frame(idx, [cur] + scene.keyframes)


def run(nt, nm):
    flags = [(p, p.requires_grad) for p in scene.decoders.parameters()]
    for p, _ in flags:
        p.requires_grad_(False)
    for _ in range(nt):
        tr._iter()
    for p, f in flags:
        p.requires_grad_(f)
    for it in range(nm):
        mp._iter('geometry' if it % 2 == 0 else 'color')


run(2, 2)                      # warm-up (lazy init, allocator)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run(n_track, n_map)
if os.environ.get('PSL_PROF_MAP', '1') != '0':      # the per-mapped-frame map updates: frustum selection + add_neural_points
    scene.map_update(src['c2w'], tr.depth, tr.color)
    select()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
