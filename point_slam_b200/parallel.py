"""Multi-GPU plumbing for the render path: one process per GPU, the neural point cloud, its
features and the decoders REPLICATED on every rank; units of work (scenes, frames, image rows) are sharded with no
data-path collective.  The only exchange is the periodic map delta after a mapping step -- appended points
(`add_neural_points`, neural_point.py:91-167), updated feature rows (`update_geo_feats/update_col_feats`,
neural_point.py:75-89) and the colour decoder (Mapper.py:368-373 optimises it) -- broadcast from the mapping rank over
NCCL (NVLink/NVSwitch); gloo is used by the CPU tests.  The delta is a few MB, i.e. latency bound: three broadcasts
(header, int64 payload, float32 payload) regardless of how many tensors changed.
"""
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) of `n_items` for `rank` (image rows, frames, scenes)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_strided(n_items: int, rank: int, world: int) -> List[int]:
    """Frames rank, rank+world, ... (keeps every rank close to the live end of the sequence)."""
    return list(range(rank, n_items, world))


@dataclass
class MapDelta:
    n_before: int                      # points before the append
    pos_new: torch.Tensor              # (K,3) f32
    geo_new: torch.Tensor              # (K,32) f32
    col_new: torch.Tensor              # (K,32) f32
    upd_idx: torch.Tensor              # (U,) i64 rows whose features changed
    geo_upd: torch.Tensor              # (U,32)
    col_upd: torch.Tensor              # (U,32)
    decoder_flat: torch.Tensor         # colour-decoder parameters, flattened in `parameters()` order


def make_delta(npc, decoders, n_before: int, upd_idx: Optional[torch.Tensor]) -> MapDelta:
    """Collect, on the mapping rank, what changed since the cloud had `n_before` points."""
    pos = npc.cloud_pos_tensor()
    geo, col = npc.get_geo_feats(), npc.get_col_feats()
    dev = geo.device
    if upd_idx is None:
        upd_idx = torch.zeros(0, dtype=torch.int64, device=dev)
    upd_idx = upd_idx.to(dev).long()
    flat = torch.cat([p.detach().reshape(-1).float() for p in decoders.color_decoder.parameters()])
    return MapDelta(n_before, pos[n_before:].float(), geo[n_before:], col[n_before:], upd_idx, geo[upd_idx], col[upd_idx], flat)


def broadcast_delta(delta: Optional[MapDelta], src: int, device, n_decoder_floats: int, group=None) -> MapDelta:
    """All ranks call this; `delta` is read on `src` and returned (as received) on every rank."""
    rank = dist.get_rank(group)
    hdr = torch.zeros(3, dtype=torch.int64, device=device)
    if rank == src:
        hdr[0], hdr[1], hdr[2] = delta.n_before, delta.pos_new.shape[0], delta.upd_idx.shape[0]
    dist.broadcast(hdr, src, group=group)
    n_before, K, U = (int(v) for v in hdr.tolist())
    n_f = K * 67 + U * 64 + n_decoder_floats
    if rank == src:
        idx = delta.upd_idx.to(device)
        pay = torch.cat([torch.cat([delta.pos_new, delta.geo_new, delta.col_new], 1).reshape(-1),
                         torch.cat([delta.geo_upd, delta.col_upd], 1).reshape(-1), delta.decoder_flat]).to(device).contiguous()
        assert pay.numel() == n_f
    else:
        idx = torch.empty(U, dtype=torch.int64, device=device)
        pay = torch.empty(n_f, dtype=torch.float32, device=device)
    if U:
        dist.broadcast(idx, src, group=group)
    dist.broadcast(pay, src, group=group)
    new = pay[:K * 67].reshape(K, 67)
    upd = pay[K * 67:K * 67 + U * 64].reshape(U, 64)
    return MapDelta(n_before, new[:, :3], new[:, 3:35], new[:, 35:], idx, upd[:, :32], upd[:, 32:], pay[K * 67 + U * 64:])


class DeltaChannel:
    """The map delta as ONE collective: a fixed-capacity float32 buffer [header | updated-row indices | new points + their
    feature rows | updated feature rows | colour decoder] broadcast with a single NCCL call per mapped frame -- no size
    negotiation round trip (the three-call `broadcast_delta` needs the header on the host before it can size the payload).
    The whole buffer travels every time (a few tens of MB over NVLink / NVSwitch: a fraction of a millisecond); the receiver
    reads the header once to slice it.  Capacity: `k_max` appended points and `u_max` updated rows per delta."""

    HDR = 4

    def __init__(self, device, n_decoder_floats: int, k_max: int = 21000, u_max: int = 1 << 17):
        self.k_max, self.u_max, self.n_dec = int(k_max), int(u_max), int(n_decoder_floats)
        self.o_idx = self.HDR
        self.o_new = self.o_idx + 2 * self.u_max                 # int64 indices viewed through float32 pairs
        self.o_upd = self.o_new + 67 * self.k_max
        self.o_dec = self.o_upd + 64 * self.u_max
        self.buf = torch.zeros(self.o_dec + self.n_dec, dtype=torch.float32, device=device)
        self._rows = torch.empty(self.u_max, 32, dtype=torch.float32, device=device)     # gather scratch (no allocator traffic per delta)
        self._hdr_host = torch.zeros(4, dtype=torch.int32).pin_memory() if torch.cuda.is_available() and str(device) != 'cpu' else torch.zeros(4, dtype=torch.int32)

    def pack(self, d: MapDelta):
        K, U = d.pos_new.shape[0], d.upd_idx.shape[0]
        assert K <= self.k_max and U <= self.u_max, 'map delta exceeds the channel capacity'
        b = self.buf
        b[:self.HDR].view(torch.int32).copy_(torch.tensor([d.n_before, K, U, 0], dtype=torch.int32))
        if U:
            b[self.o_idx:self.o_idx + 2 * U].view(torch.int64).copy_(d.upd_idx)
            b[self.o_upd:self.o_upd + 64 * U].view(U, 64).copy_(torch.cat([d.geo_upd, d.col_upd], 1))
        if K:
            b[self.o_new:self.o_new + 67 * K].view(K, 67).copy_(torch.cat([d.pos_new, d.geo_new, d.col_new], 1))
        b[self.o_dec:].copy_(d.decoder_flat)

    def pack_from(self, npc, decoders, n_before: int, upd_idx: Optional[torch.Tensor]):
        """make_delta + pack without temporaries: the changed rows are gathered straight into the channel buffer."""
        pos, geo, col = npc.cloud_pos_tensor(), npc.get_geo_feats(), npc.get_col_feats()
        K = pos.shape[0] - int(n_before)
        U = 0 if upd_idx is None else int(upd_idx.shape[0])
        assert 0 <= K <= self.k_max and U <= self.u_max, 'map delta exceeds the channel capacity'
        b = self.buf
        self._hdr_host[0], self._hdr_host[1], self._hdr_host[2] = int(n_before), K, U
        b[:self.HDR].view(torch.int32).copy_(self._hdr_host, non_blocking=True)
        if K:
            new = b[self.o_new:self.o_new + 67 * K].view(K, 67)
            new[:, :3].copy_(pos[n_before:]); new[:, 3:35].copy_(geo[n_before:]); new[:, 35:].copy_(col[n_before:])
        if U:
            idx = upd_idx.to(b.device).long()
            b[self.o_idx:self.o_idx + 2 * U].view(torch.int64).copy_(idx)
            upd = b[self.o_upd:self.o_upd + 64 * U].view(U, 64)
            torch.index_select(geo, 0, idx, out=self._rows[:U]); upd[:, :32].copy_(self._rows[:U])
            torch.index_select(col, 0, idx, out=self._rows[:U]); upd[:, 32:].copy_(self._rows[:U])
        off = self.o_dec
        for p in decoders.color_decoder.parameters():
            b[off:off + p.numel()].copy_(p.detach().reshape(-1))
            off += p.numel()

    def unpack(self) -> MapDelta:
        b = self.buf
        n_before, K, U, _ = (int(v) for v in b[:self.HDR].view(torch.int32).tolist())     # the receiver's one host read
        new = b[self.o_new:self.o_new + 67 * K].view(K, 67)
        upd = b[self.o_upd:self.o_upd + 64 * U].view(U, 64)
        idx = b[self.o_idx:self.o_idx + 2 * U].view(torch.int64)
        return MapDelta(n_before, new[:, :3], new[:, 3:35], new[:, 35:], idx, upd[:, :32], upd[:, 32:], b[self.o_dec:])

    def broadcast(self, delta: Optional[MapDelta], src: int, group=None) -> MapDelta:
        """All ranks call this; `delta` is read on `src` (None: the source already filled the buffer with pack_from).  One dist.broadcast."""
        if dist.get_rank(group) == src and delta is not None:
            self.pack(delta)
        dist.broadcast(self.buf, src, group=group)
        return self.unpack()


def apply_delta(npc, decoders, d: MapDelta):
    """Bring a replica up to date (no-op data-wise on the source rank, but cheap enough to run everywhere)."""
    assert npc.pts_num() in (d.n_before, d.n_before + d.pos_new.shape[0]), 'replica diverged from the mapping rank'
    if npc.pts_num() == d.n_before and d.pos_new.shape[0]:
        npc.append_points(d.pos_new, d.geo_new, d.col_new)
    if d.upd_idx.numel():
        npc.update_geo_feats(d.geo_upd, d.upd_idx)
        npc.update_col_feats(d.col_upd, d.upd_idx)
    off = 0
    with torch.no_grad():
        for p in decoders.color_decoder.parameters():
            n = p.numel()
            p.copy_(d.decoder_flat[off:off + n].reshape(p.shape))
            off += n


def n_decoder_floats(decoders) -> int:
    return sum(p.numel() for p in decoders.color_decoder.parameters())


def render_img_sharded(render_rows, H: int, W: int, device, group=None):
    """Row-sharded full-image render (strong scaling of Renderer.render_img, Renderer.py:204-283).
    `render_rows(r0, r1)` -> depth (r1-r0, W), uncertainty (r1-r0, W), color (r1-r0, W, 3) for image rows [r0, r1).
    Every rank returns the full (H,W) images; the exchange is one all_gather of 5 floats per pixel."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    r0, r1 = shard_range(H, rank, world)
    rows_max = (H + world - 1) // world
    buf = torch.zeros(rows_max, W, 5, dtype=torch.float32, device=device)
    if r1 > r0:
        d, u, c = render_rows(r0, r1)
        buf[:r1 - r0, :, 0], buf[:r1 - r0, :, 1], buf[:r1 - r0, :, 2:] = d.float(), u.float(), c.float()
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(H, r, world)
        parts.append(out[r][:b - a])
    full = torch.cat(parts, 0)
    return full[..., 0], full[..., 1], full[..., 2:]


def rerender_frames(renderer, npc, decoders, frames, device, every_frame=1, group=None):
    """Frame-parallel end-of-run re-render (Mapper.py:826-876: every `every_frame`-th frame of the sequence is rendered from its
    estimated pose and compared with the sensor images).  frames[k] = dict(c2w (4,4) / (3,4), depth (H,W), color (H,W,3), dyn_r_query
    (H,W) f64 or None) on the device; rank r of a process group renders frames r, r + world, ... of the selected ones (no
    collective until the end: one all_reduce of three sums).  -> dict(frames, psnr, depth_l1) averaged over all ranks' frames,
    with the reference's definitions: PSNR = -10 log10(mse over pixels with depth > 0) (:868-870), depth L1 = mean |d_gt - d| over
    the same pixels (:884-885).  (MS-SSIM / LPIPS need pytorch_msssim / torchmetrics, which this image lacks.)"""
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_available() and dist.is_initialized() else (0, 1)
    sel = list(range(0, len(frames), every_frame))
    acc = torch.zeros(3, dtype=torch.float64, device=device)
    outs = {}
    for k in shard_strided(len(sel), rank, world):
        f = frames[sel[k]]
        d, _, c = renderer.render_img(npc, decoders, f['c2w'], device, 'color', gt_depth=f['depth'], npc_geo_feats=npc.get_geo_feats(),
                                      npc_col_feats=npc.get_col_feats(), dynamic_r_query=f.get('dyn_r_query'),
                                      cloud_pos=npc.cloud_pos_tensor())
        # masked means as weighted sums: boolean-mask indexing would cost four device->host syncs (nonzero) per frame and leave
        # the GPU idle while the host launches the next frame
        m = (f['depth'] > 0)
        cnt = m.sum().clamp(min=1).double()
        mse = (((f['color'] - c.float()) ** 2).sum(-1) * m).sum().double() / (3.0 * cnt)
        l1 = (torch.abs(f['depth'] - d.float()) * m).sum().double() / cnt
        acc += torch.stack([-10.0 * torch.log10(mse), l1, torch.ones((), dtype=torch.float64, device=device)])
        outs[sel[k]] = (d, c)
    if world > 1:
        dist.all_reduce(acc, group=group)
    n = max(float(acc[2]), 1.0)
    return dict(frames=int(acc[2]), psnr=float(acc[0]) / n, depth_l1=float(acc[1]) / n, rendered=outs)
