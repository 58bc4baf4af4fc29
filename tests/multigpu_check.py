"""torchrun entry (2+ ranks, NCCL): replicated map stays identical after a map-delta broadcast from the mapping rank, and
a row-sharded render_img equals the single-GPU image.  Run by tests/test_gpu_scale.py when >= 2 GPUs are visible."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from point_slam_b200 import parallel as PL, synth  # noqa: E402
from tests.test_gpu_scale import _scene  # noqa: E402
import tests.test_gpu_scale as T  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'
    T.DEV = dev
    dist.init_process_group('nccl', device_id=torch.device(dev))
    cfg, dec, npc, ren, cloud = _scene(200_000)
    nd = PL.n_decoder_floats(dec)
    delta = None
    if rank == 0:
        n0 = npc.pts_num()
        pose = synth.trajectory(3)[0]
        depth, color = synth.make_frame(pose)
        o, d = synth.pixel_rays(pose, 480, 640, 517.3, 516.5, 318.6, 255.3)
        pix = np.random.default_rng(2).integers(0, 480 * 640, 4000)
        ro = torch.from_numpy(np.broadcast_to(o, (4000, 3)).astype(np.float32).copy()).to(dev)
        rd = torch.from_numpy(d.reshape(-1, 3)[pix].astype(np.float32)).to(dev)
        gd = torch.from_numpy(depth.reshape(-1)[pix]).to(dev)
        npc.add_neural_points(ro, rd, gd, torch.zeros(4000, 3, device=dev), is_pts_grad=True)      # radius_min: some get added
        idx = torch.arange(100, 5000, 7, device=dev)
        npc.update_geo_feats(npc.get_geo_feats()[idx] + 0.5, idx)
        npc.update_col_feats(npc.get_col_feats()[idx] - 0.5, idx)
        with torch.no_grad():
            for p in dec.color_decoder.parameters():
                p.mul_(1.01)
        delta = PL.make_delta(npc, dec, n0, idx)
    got = PL.broadcast_delta(delta, 0, dev, nd)
    PL.apply_delta(npc, dec, got)
    flat = torch.cat([npc.cloud_pos_tensor().reshape(-1), npc.get_geo_feats().reshape(-1), npc.get_col_feats().reshape(-1),
                      torch.cat([p.detach().reshape(-1) for p in dec.color_decoder.parameters()])]).double()
    sig = torch.stack([flat.sum(), flat.abs().sum(), torch.tensor(float(npc.pts_num()), device=dev, dtype=torch.float64)])
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    assert all(torch.equal(s, sigs[0]) for s in sigs), f'replicas differ: {sigs}'
    assert npc.index_ntotal() == npc.pts_num()
    # row-sharded full image against the replicated (updated) map == single-GPU image
    H, W = 96, 128
    ren.H, ren.W, ren.fx, ren.fy, ren.cx, ren.cy = H, W, 103.46, 103.3, 63.72, 51.06
    pose = synth.trajectory(3)[1]
    depth, _ = synth.make_frame(pose, dict(H=H, W=W, fx=ren.fx, fy=ren.fy, cx=ren.cx, cy=ren.cy))
    gd = torch.from_numpy(depth).to(dev)
    dyn = torch.full((H, W), 0.1, dtype=torch.float64, device=dev)
    c2w = torch.from_numpy(pose[:3, :4].astype(np.float32)).to(dev)
    fixed = (torch.zeros(32, device=dev), torch.zeros(32, device=dev))
    dec.draw_no_neighbor_vectors = lambda stage, device: fixed
    from point_slam_b200.src import common
    ro, rd = common.get_rays(H, W, ren.fx, ren.fy, ren.cx, ren.cy, c2w, dev)

    def render_rows(r0, r1):
        with torch.no_grad():
            d_, u_, c_, _ = ren.render_batch_ray(npc, dec, rd[r0:r1].reshape(-1, 3), ro[r0:r1].reshape(-1, 3), dev, 'color',
                                                 gt_depth=gd[r0:r1].reshape(-1), npc_geo_feats=npc.get_geo_feats(),
                                                 npc_col_feats=npc.get_col_feats(), cloud_pos=npc.cloud_pos_tensor(),
                                                 dynamic_r_query=dyn[r0:r1].reshape(-1))
        return d_.reshape(r1 - r0, W), u_.reshape(r1 - r0, W), c_.reshape(r1 - r0, W, 3)

    d_sh, u_sh, c_sh = PL.render_img_sharded(render_rows, H, W, dev)
    d1, u1, c1 = ren.render_img(npc, dec, c2w, dev, 'color', gt_depth=gd, npc_geo_feats=npc.get_geo_feats(),
                                npc_col_feats=npc.get_col_feats(), dynamic_r_query=dyn, cloud_pos=npc.cloud_pos_tensor())
    assert torch.equal(d_sh, d1.float()) and torch.equal(c_sh, c1), 'row-sharded image differs from the single-GPU image'
    dist.barrier()
    if rank == 0:
        print('MULTIGPU OK', world)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
