"""world_size-2 gloo tests of the multi-GPU host logic (map-delta broadcast, sharding, row-sharded image gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from point_slam_b200 import parallel as PL


class FakeNpc:
    """CPU stand-in with the NeuralPointCloud methods parallel.py uses (the real class needs a GPU for its hash)."""

    def __init__(self, n, seed):
        g = torch.Generator().manual_seed(seed)
        self._pos = torch.randn(n, 3, generator=g)
        self.geo_feats = torch.randn(n, 32, generator=g)
        self.col_feats = torch.randn(n, 32, generator=g)
        self.rebuilds = 0

    def pts_num(self):
        return self._pos.shape[0]

    def cloud_pos_tensor(self):
        return self._pos

    def get_geo_feats(self):
        return self.geo_feats

    def get_col_feats(self):
        return self.col_feats

    def update_geo_feats(self, feats, indices=None):
        self.geo_feats[indices] = feats.detach().clone()

    def update_col_feats(self, feats, indices=None):
        self.col_feats[indices] = feats.detach().clone()

    def append_points(self, pts, g, c):
        self._pos = torch.cat([self._pos, pts], 0)
        self.geo_feats = torch.cat([self.geo_feats, g], 0)
        self.col_feats = torch.cat([self.col_feats, c], 0)
        self.rebuilds += 1


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from point_slam_b200.default_config import make_cfg
        from point_slam_b200.src.conv_onet import config as model_config
        torch.manual_seed(1219)
        dec = model_config.get_model(make_cfg('replica', 'cpu'))
        npc = FakeNpc(1000, seed=3)                       # identical replicas to start with
        nd = PL.n_decoder_floats(dec)
        delta = None
        if rank == 0:                                     # the mapping rank mutates its copy
            n0 = npc.pts_num()
            g = torch.Generator().manual_seed(11)
            npc.append_points(torch.randn(30, 3, generator=g), torch.randn(30, 32, generator=g), torch.randn(30, 32, generator=g))
            idx = torch.tensor([5, 17, 999, 1003])
            npc.update_geo_feats(torch.randn(4, 32, generator=g), idx)
            npc.update_col_feats(torch.randn(4, 32, generator=g), idx)
            with torch.no_grad():
                for p in dec.color_decoder.parameters():
                    p.add_(0.01)
            delta = PL.make_delta(npc, dec, n0, idx)
        got = PL.broadcast_delta(delta, 0, 'cpu', nd)
        PL.apply_delta(npc, dec, got)
        # every rank must now hold identical state: compare checksums across ranks
        flat = torch.cat([npc.cloud_pos_tensor().reshape(-1), npc.get_geo_feats().reshape(-1), npc.get_col_feats().reshape(-1),
                          torch.cat([p.detach().reshape(-1) for p in dec.color_decoder.parameters()])]).double()
        sig = torch.stack([flat.sum(), (flat * torch.arange(flat.numel(), dtype=torch.float64)).sum(), torch.tensor(float(npc.pts_num()))])
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        assert all(torch.equal(s, sigs[0]) for s in sigs), 'replicas differ after the delta broadcast'
        assert npc.pts_num() == 1030
        # the same exchange as ONE collective (fixed-capacity DeltaChannel): a second delta on top of the first
        ch = PL.DeltaChannel('cpu', nd, k_max=64, u_max=16)
        delta = None
        if rank == 0:
            n0 = npc.pts_num()
            g = torch.Generator().manual_seed(12)
            npc.append_points(torch.randn(7, 3, generator=g), torch.randn(7, 32, generator=g), torch.randn(7, 32, generator=g))
            idx = torch.tensor([0, 1031, 12])
            npc.update_geo_feats(torch.randn(3, 32, generator=g), idx)
            with torch.no_grad():
                for p in dec.color_decoder.parameters():
                    p.mul_(1.01)
            ch.pack_from(npc, dec, n0, idx)                # gathers straight into the channel buffer
        PL.apply_delta(npc, dec, ch.broadcast(None, 0))
        flat = torch.cat([npc.cloud_pos_tensor().reshape(-1), npc.get_geo_feats().reshape(-1), npc.get_col_feats().reshape(-1),
                          torch.cat([p.detach().reshape(-1) for p in dec.color_decoder.parameters()])]).double()
        sig = torch.stack([flat.sum(), (flat * torch.arange(flat.numel(), dtype=torch.float64)).sum(), torch.tensor(float(npc.pts_num()))])
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        assert all(torch.equal(s, sigs[0]) for s in sigs), 'replicas differ after the single-collective delta'
        assert npc.pts_num() == 1037
        # row-sharded image: each rank renders its rows of a synthetic "image"
        H, W = 11, 7
        def render_rows(r0, r1):
            rows = torch.arange(r0, r1, dtype=torch.float32)[:, None].expand(r1 - r0, W)
            return rows, rows * 2, torch.stack([rows, rows + 1, rows + 2], -1)
        d, u, c = PL.render_img_sharded(render_rows, H, W, 'cpu')
        want = torch.arange(H, dtype=torch.float32)[:, None].expand(H, W)
        assert torch.equal(d, want) and torch.equal(u, want * 2) and torch.equal(c[..., 2], want + 2)
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_map_delta_broadcast_and_sharded_image_gloo():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) == 1 and ret.get(1) == 1


def test_sharding_helpers():
    for n in (0, 1, 7, 8, 480):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = PL.shard_range(n, r, world)
                assert 0 <= a <= b <= n
                seen += list(range(a, b))
            assert seen == list(range(n))
            strided = sorted(sum((PL.shard_strided(n, r, world) for r in range(world)), []))
            assert strided == list(range(n))
            sizes = [PL.shard_range(n, r, world)[1] - PL.shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_rerender_metrics_match_masked_selection():
    """rerender_frames computes PSNR / depth L1 over the pixels with sensor depth > 0 (Mapper.py:868-885) as weighted sums (no
    boolean-mask indexing, i.e. no device->host sync per frame): same numbers as the selection-based definition."""
    import math
    import types
    g = torch.Generator().manual_seed(3)
    H, W = 24, 32
    frames, truth = [], []
    for k in range(3):
        depth = torch.rand(H, W, generator=g) * 3
        depth[torch.rand(H, W, generator=g) < 0.2] = 0.0
        color = torch.rand(H, W, 3, generator=g)
        rd, rc = depth + 0.05 * torch.randn(H, W, generator=g), (color + 0.03 * torch.randn(H, W, 3, generator=g)).clamp(0, 1)
        frames.append(dict(c2w=torch.eye(4), depth=depth, color=color, dyn_r_query=None, _out=(rd, rc)))
        m = depth > 0
        truth.append((-10.0 * math.log10(float(torch.nn.functional.mse_loss(color[m], rc[m]))), float((depth[m] - rd[m]).abs().mean())))
    it = iter(frames)
    renderer = types.SimpleNamespace(render_img=lambda *a, **kw: (lambda f: (f['_out'][0], None, f['_out'][1]))(next(it)))
    npc = types.SimpleNamespace(get_geo_feats=lambda: None, get_col_feats=lambda: None, cloud_pos_tensor=lambda: None)
    out = PL.rerender_frames(renderer, npc, None, frames, 'cpu')
    assert out['frames'] == 3
    assert abs(out['psnr'] - sum(t[0] for t in truth) / 3) < 1e-4
    assert abs(out['depth_l1'] - sum(t[1] for t in truth) / 3) < 1e-6
