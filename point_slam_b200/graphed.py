"""Launch-bound inner loops as CUDA graphs ("next #2").

The reference's optimisation iterations (Tracker.optimize_cam_in_batch, the joint loop of Mapper.optimize_map) issue ~450
tiny ATen kernels and ~10 stream synchronisations each, because rays are compacted with boolean masks (data-dependent
shapes).  Here the same iteration is written with STATIC shapes -- every sampled pixel stays in the batch, rays that the
reference would have dropped (no depth, depth outlier) carry depth 0 and zero loss weight -- so that one whole iteration
(pixel sampling on the device RNG, ray generation, fused render forward, loss, backward, Adam) is captured once into a
CUDA graph and replayed with a single launch.  Sums over the kept rays are identical to the reference's sums over the
compacted rays; the dropped rays cost a few percent of extra render work.
"""
import torch

from .src import common


def _masked_stats(depth, valid):
    """10*median and 1.2*max of the valid depths (Tracker.py:142-143 / Mapper.py:507-509) without compaction."""
    nan = torch.full_like(depth, float('nan'))
    med = torch.nanmedian(torch.where(valid, depth, nan))
    mx = torch.max(torch.where(valid, depth, torch.full_like(depth, -float('inf'))))
    return torch.minimum(10 * med, 1.2 * mx)


def _msum(x, mask):
    return torch.where(mask, x, torch.zeros_like(x)).sum()


def sample_uv_device(H0, H1, W0, W1, n, depth, color, device):
    """common.get_sample_uv (common.py:77-89) without host-side meshgrids: the same single torch.randint draw over the
    flattened window, pixel coordinates recovered arithmetically (identical values: the meshgrid holds exact integers)."""
    ww = W1 - W0
    idx = torch.randint((H1 - H0) * ww, (n,), device=device)
    jj = torch.div(idx, ww, rounding_mode='floor') + H0
    ii = idx - (jj - H0) * ww + W0
    return ii.float(), jj.float(), depth[jj, ii], color[jj, ii]


def tracker_iteration_static(renderer, npc, decoders, cam, gt_color, gt_depth, dyn_r_query, intr, n_pixels, device,
                             geo_feats, col_feats, cloud_pos, edge, w_color=0.5):
    """One tracking iteration, static shapes, no host synchronisation.  Returns the loss (backward already run)."""
    H, W = intr['H'], intr['W']
    c2w = common.get_camera_from_tensor(cam)
    i, j, b_depth, b_color = sample_uv_device(edge[0], H - edge[0], edge[1], W - edge[1], n_pixels, gt_depth, gt_color, device)
    rays_o, rays_d = common.get_rays_from_uv(i, j, c2w, intr['fx'], intr['fy'], intr['cx'], intr['cy'], device)
    b_rq = dyn_r_query[j.long(), i.long()] if dyn_r_query is not None else None
    with torch.no_grad():
        valid = b_depth > 0
        inside = valid & (b_depth <= _masked_stats(b_depth, valid))
        depth_in = torch.where(inside, b_depth, torch.zeros_like(b_depth))
    depth, unc, color, _ = renderer.render_batch_ray(npc, decoders, rays_d, rays_o.contiguous(), device, stage='color',
                                                     gt_depth=depth_in, npc_geo_feats=geo_feats, npc_col_feats=col_feats,
                                                     is_tracker=True, cloud_pos=cloud_pos, dynamic_r_query=b_rq,
                                                     _zero_depth=(None, None))
    unc = unc.detach()
    tmp = torch.abs(depth_in - depth) / torch.sqrt(unc + 1e-10)
    with torch.no_grad():
        ok = inside & (~torch.isnan(depth)) & (~torch.isnan(unc))
        mean_tmp = _msum(tmp, inside) / inside.sum().clamp_min(1)
        mask = ok & (tmp < 10 * mean_tmp)
    loss = _msum(torch.clamp(tmp, min=0.0, max=1e3), mask) + w_color * _msum(torch.abs(b_color - color), mask[:, None].expand(-1, 3))
    loss.backward()
    return loss.detach()


def stack_keyframes(keyframes):
    """list of keyframe dicts -> one dict of stacked device tensors (done once per mapped frame)."""
    out = dict(color=torch.stack([k['color'] for k in keyframes]), depth=torch.stack([k['depth'] for k in keyframes]),
               c2w=torch.stack([k['c2w'][:3, :4] for k in keyframes]), dyn_r_query=None)
    if keyframes[0].get('dyn_r_query') is not None:
        out['dyn_r_query'] = torch.stack([k['dyn_r_query'] for k in keyframes])
    return out


def mapper_iteration_static(renderer, npc, decoders, state, kfs, intr, n_pixels, device, stage, cloud_pos, w_color=0.1):
    """One mapping iteration over the stacked keyframes `kfs` (see stack_keyframes): static shapes, no host sync.
    All keyframes are sampled with one batched draw (the reference loops over them, Mapper.py:459-500)."""
    H, W = intr['H'], intr['W']
    idx = state.indices
    npc_geo, npc_col = state.npc_geo, state.npc_col
    npc_geo[idx] = state.geo
    npc_col[idx] = state.col
    K = kfs['depth'].shape[0]
    per = n_pixels // K
    pix = torch.randint(H * W, (K, per), device=device)
    jj = torch.div(pix, W, rounding_mode='floor')
    ii = pix - jj * W
    kk = torch.arange(K, device=device)[:, None].expand(K, per)
    b_depth = kfs['depth'][kk, jj, ii].reshape(-1)
    b_color = kfs['color'][kk, jj, ii].reshape(-1, 3)
    b_rq = kfs['dyn_r_query'][kk, jj, ii].reshape(-1) if kfs['dyn_r_query'] is not None else None
    dirs = torch.stack([(ii.float() - intr['cx']) / intr['fx'], -(jj.float() - intr['cy']) / intr['fy'],
                        -torch.ones(K, per, device=device)], -1)                       # common.py:49-50
    c2w = kfs['c2w']
    rays_d = torch.sum(dirs[..., None, :] * c2w[:, None, :3, :3], -1).reshape(-1, 3)   # common.py:53
    rays_o = c2w[:, None, :3, 3].expand(K, per, 3).reshape(-1, 3)
    with torch.no_grad():
        valid = b_depth > 0
        inside = valid & (b_depth <= _masked_stats(b_depth, valid))
        depth_in = torch.where(inside, b_depth, torch.zeros_like(b_depth))
    depth, unc, color, vmask = renderer.render_batch_ray(npc, decoders, rays_d, rays_o, device, stage, gt_depth=depth_in,
                                                         npc_geo_feats=npc_geo, npc_col_feats=npc_col, is_tracker=False,
                                                         cloud_pos=cloud_pos, dynamic_r_query=b_rq, _zero_depth=(None, None))
    m = inside & vmask & (~torch.isnan(depth))
    loss = _msum(torch.abs(depth_in - depth), m)
    if stage == 'color':
        loss = loss + w_color * _msum(torch.abs(b_color - color), m[:, None].expand(-1, 3))
    loss.backward()
    return loss.detach()


def _reset_adam(opt):
    """Fresh-optimizer semantics (the reference builds a new Adam per frame) without reallocating capturable state."""
    for st in opt.state.values():
        for k, v in st.items():
            if torch.is_tensor(v):
                v.zero_()


class GraphedTracker:
    """Tracker.optimize_cam_in_batch x n_iters as replays of one captured CUDA graph."""

    def __init__(self, renderer, npc, decoders, intr, n_pixels, device, edge=(20, 20), lr=0.002, w_color=0.5):
        self.r, self.npc, self.dec, self.intr, self.n, self.dev, self.edge, self.w = renderer, npc, decoders, intr, n_pixels, device, edge, w_color
        H, W = intr['H'], intr['W']
        self.color = torch.zeros(H, W, 3, device=device)
        self.depth = torch.zeros(H, W, device=device)
        self.dyn = torch.zeros(H, W, dtype=torch.float64, device=device)
        self.cam = torch.zeros(7, device=device, requires_grad=True)
        self.opt = torch.optim.Adam([self.cam], lr=lr, capturable=True, fused=True)
        self.loss = torch.zeros((), device=device)
        self.graph = None

    def _iter(self):
        # the pose is the only optimised quantity: decoder/feature gradients (which the reference's autograd computes and
        # throws away) are not requested, so the backward kernel skips every weight-gradient phase
        self.opt.zero_grad(set_to_none=True)
        loss = tracker_iteration_static(self.r, self.npc, self.dec, self.cam, self.color, self.depth, self.dyn, self.intr, self.n,
                                        self.dev, self.npc.get_geo_feats(), self.npc.get_col_feats(), self.npc.cloud_pos_tensor(),
                                        self.edge, self.w)
        self.opt.step()
        self.loss.copy_(loss)

    def load_frame(self, color, depth, dyn, cam_init):
        self.color.copy_(color, non_blocking=True); self.depth.copy_(depth, non_blocking=True); self.dyn.copy_(dyn, non_blocking=True)
        with torch.no_grad():
            self.cam.copy_(cam_init)
        _reset_adam(self.opt)

    def capture(self):
        flags = [(p, p.requires_grad) for p in self.dec.parameters()]
        for p, _ in flags:
            p.requires_grad_(False)
        try:
            self._capture()
        finally:
            for p, f in flags:
                p.requires_grad_(f)

    def _capture(self):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        cam0 = self.cam.detach().clone()
        with torch.cuda.stream(s):
            for _ in range(2):
                self._iter()
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iter()
        with torch.no_grad():
            self.cam.copy_(cam0)
        _reset_adam(self.opt)

    def run(self, n_iters):
        if self.graph is None:
            self.capture()
        for _ in range(n_iters):
            self.graph.replay()
        return self.loss


class _MapState:
    pass


class GraphedMapper:
    """The joint loop of Mapper.optimize_map as replays of ONE captured graph per stage for the whole run.

    The frustum-selected feature slices change size from frame to frame; to keep every shape static the index list is
    padded to a fixed capacity `u_max` with the index of a dummy extra feature row (row N of (N+1)-row copies of the feature
    tensors): padded slots scatter into / gather from that row, which no kNN index ever references, so they receive zero
    gradient and Adam leaves them at zero.  Graphs are re-captured only when the cloud size, the capacity or the spatial
    hash changes (e.g. after add_neural_points)."""

    def __init__(self, renderer, npc, decoders, intr, n_pixels, device, w_color=0.1, lr_dec=0.005, lr_geo=0.005, lr_col=0.005,
                 u_max=1 << 17):
        self.r, self.npc, self.dec, self.intr, self.n, self.dev, self.w = renderer, npc, decoders, intr, n_pixels, device, w_color
        self.lrs = (lr_dec, lr_geo, lr_col)
        self.loss = torch.zeros((), device=device)
        self.u_max = int(u_max)
        self.graphs = {}
        self.state = None
        self.key = None
        self.warm = False

    def _alloc(self, N, n_kf):
        H, W = self.intr['H'], self.intr['W']
        d = self.dev
        st = _MapState()
        st.indices = torch.full((self.u_max,), N, dtype=torch.int64, device=d)
        st.npc_geo = torch.zeros(N + 1, 32, device=d)
        st.npc_col = torch.zeros(N + 1, 32, device=d)
        st.geo = torch.zeros(self.u_max, 32, device=d, requires_grad=True)
        st.col = torch.zeros(self.u_max, 32, device=d, requires_grad=True)
        st.optimizer = torch.optim.Adam([{'params': list(self.dec.color_decoder.parameters()), 'lr': self.lrs[0]},
                                         {'params': [st.geo], 'lr': self.lrs[1]}, {'params': [st.col], 'lr': self.lrs[2]}],
                                        capturable=True, fused=True)
        self.state = st
        self.keyframes = dict(color=torch.zeros(n_kf, H, W, 3, device=d), depth=torch.zeros(n_kf, H, W, device=d),
                              c2w=torch.zeros(n_kf, 3, 4, device=d), dyn_r_query=torch.zeros(n_kf, H, W, dtype=torch.float64, device=d))
        self.graphs = {}

    def begin_frame(self, indices, keyframes):
        """indices: (U,) int64 rows of the feature tensors to optimise; keyframes: list of dicts (color, depth, c2w, dyn_r_query)."""
        N, U = self.npc.pts_num(), int(indices.shape[0])
        while U > self.u_max:
            self.u_max *= 2
        key = (N, self.u_max, len(keyframes), self.npc.spatial_hash().sorted_pts.data_ptr())
        if key != self.key:
            self._alloc(N, len(keyframes))
            self.key = key
        st = self.state
        self.n_used = U
        with torch.no_grad():
            st.indices.fill_(N)
            st.indices[:U] = indices
            st.npc_geo[:N] = self.npc.get_geo_feats()
            st.npc_col[:N] = self.npc.get_col_feats()
            st.geo.copy_(st.npc_geo[st.indices])
            st.col.copy_(st.npc_col[st.indices])
            for i, kf in enumerate(keyframes):
                self.keyframes['color'][i].copy_(kf['color']); self.keyframes['depth'][i].copy_(kf['depth'])
                self.keyframes['c2w'][i].copy_(kf['c2w'][:3, :4]); self.keyframes['dyn_r_query'][i].copy_(kf['dyn_r_query'])
        _reset_adam(st.optimizer)

    def _iter(self, stage):
        st = self.state
        st.optimizer.zero_grad(set_to_none=True)
        loss = mapper_iteration_static(self.r, self.npc, self.dec, st, self.keyframes, self.intr, self.n, self.dev, stage,
                                       self.npc.cloud_pos_tensor(), self.w)
        st.optimizer.step()
        st.npc_geo, st.npc_col = st.npc_geo.detach(), st.npc_col.detach()
        self.loss.copy_(loss)

    def run(self, stage, n_iters):
        done = 0
        if stage not in self.graphs:
            if not self.warm:                              # one-time lazy initialisation outside of capture (executes once)
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._iter(stage)
                torch.cuda.current_stream().wait_stream(s)
                self.warm = True
                done = 1
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):                      # capture records the iteration, it does not execute it
                self._iter(stage)
            self.graphs[stage] = g
        for _ in range(n_iters - done):
            self.graphs[stage].replay()
        return self.loss

    def write_back(self):
        """Optimised slices -> the neural point cloud (Mapper.py:605-610)."""
        st, U = self.state, self.n_used
        self.npc.update_geo_feats(st.geo.detach()[:U], st.indices[:U])
        self.npc.update_col_feats(st.col.detach()[:U], st.indices[:U])
