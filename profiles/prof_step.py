"""Profiling driver: the bench scene, a few tracker / mapper iterations inside a cudaProfilerStart/Stop window.

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
scene = bench.GpuScene(0, dev, 500000, 1)
cur = scene.resident[0]
npc, dec, render = scene.npc, scene.decoders, scene.renderer.render_batch_ray
cloud = npc.cloud_pos_tensor()
cam = bench.cam_tensor_from_c2w(scene.frames_host[bench.N_KEYFRAMES]['c2w'], 0.01, scene.rng).to(dev).requires_grad_(True)
opt = torch.optim.Adam([cam], lr=0.002)
idx = IT.frustum_indices(cloud, cur['c2w'], bench.INTR)
state = IT.MapperState(npc, dec, idx)
kfs = [cur] + scene.keyframes


def run(nt, nm):
    for _ in range(nt):
        IT.tracker_iteration(render, npc, dec, cam, opt, cur['color'], cur['depth'], cur['dyn_r_query'], bench.INTR,
                             bench.TRACK_PIX, dev, npc.get_geo_feats(), npc.get_col_feats(), cloud, edge=(100, 100))
    for it in range(nm):
        IT.mapper_iteration(render, npc, dec, state, kfs, bench.INTR, bench.MAP_PIX, dev,
                            'geometry' if it % 2 == 0 else 'color', cloud)


run(2, 2)                      # warm-up (lazy init, allocator)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run(n_track, n_map)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
