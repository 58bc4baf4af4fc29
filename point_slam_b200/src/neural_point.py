"""Drop-in for the reference `src/neural_point.py`: same class name, constructor and the 14 public methods
(neural_point.py:44-277), backed by a GPU-resident spatial hash and the exact radius-kNN kernel instead of a
faiss GpuIndexIVFFlat.  Positions and the two feature tensors live in capacity-doubling device buffers (the reference
re-allocates all of them with list `+=` / `torch.cat` on every add, neural_point.py:147-159); the list views the reference
API promises (`cloud_pos()`, `_cloud_pos`, `input_pos()`, `input_rgb()`) are materialised lazily.  `add_neural_points` is
one library call (`ops.add_points`: depth mask, add-radius test, ordered compaction, point emission) plus one 8-byte
device->host read of the counts.
"""
import numpy as np
import torch

from .. import ops
from .common import setup_seed


class _IndexShim:
    """Duck-type of the faiss index attributes offline tools poke at (get_mesh_tsdf_fusion.py:66-82)."""

    def __init__(self, owner):
        self._o = owner
        self.nprobe = owner.cfg['pointcloud'].get('nprobe', 1)    # accepted and ignored: the search is exact
        self.is_trained = False

    @property
    def ntotal(self):
        return self._o._indexed

    def train(self, xb):
        self.is_trained = True

    def add(self, xb):
        self._o._index_add(xb)

    def search(self, q, k):
        assert k == self._o.nn_num
        D, I = self._o._search_all(q)
        return D, I


class NeuralPointCloud(object):
    def __init__(self, cfg):
        self.cfg = cfg
        self.c_dim = cfg['model']['c_dim']
        self.device = cfg['mapping']['device']
        self.cuda_id = 0
        self.use_dynamic_radius = cfg['use_dynamic_radius']
        pc = cfg['pointcloud']
        self.nn_num = pc['nn_num']
        assert self.nn_num == 8, 'kernels are specialised for nn_num = 8'
        self.nlist = pc['nlist']
        self.radius_add, self.radius_min, self.radius_query = pc['radius_add'], pc['radius_min'], pc['radius_query']
        self.fix_interval_when_add_along_ray = pc['fix_interval_when_add_along_ray']
        self.N_surface = cfg['rendering']['N_surface']
        self.N_add = pc['N_add']
        self.near_end_surface, self.far_end_surface = pc['near_end_surface'], pc['far_end_surface']

        self._pos_buf = torch.zeros((0, 3), dtype=torch.float32, device=self.device)   # capacity-doubling storage
        self._pos = self._pos_buf[:0]                                                  # all positions ever appended (view)
        self._pos_list_cache = None
        self._geo_buf = self._col_buf = None
        self._input_pos_list, self._input_rgb_list = [], []
        self._input_pending = []           # device chunks (pos, rgb) not yet converted to the host lists
        self._pts_num = 0
        self._indexed = 0                  # number of positions the hash currently covers
        self._store_gen = 0                # bumped when a position / feature buffer is re-allocated
        self.geo_feats = None
        self.col_feats = None
        self.keyframe_dict = []
        self._grid = ops.SpatialHash(cell=float(pc.get('hash_cell', self.radius_query)))
        self.index = _IndexShim(self)
        setup_seed(cfg['setup_seed'])

    # ---- storage accessors (neural_point.py:44-89) ----------------------------------------------------------------
    @property
    def _cloud_pos(self):
        if self._pos_list_cache is None:
            self._pos_list_cache = self._pos.tolist()
        return self._pos_list_cache

    @_cloud_pos.setter
    def _cloud_pos(self, value):          # offline tools assign a list (get_mesh_tsdf_fusion.py:66)
        self._pos = torch.as_tensor(value, dtype=torch.float32, device=self.device).reshape(-1, 3)   # adopted by _reserve
        self._pos_list_cache = None
        self._indexed = 0                  # the hash covers nothing of the new positions until index.add() is called

    def cloud_pos(self, index=None):
        return self._cloud_pos if index is None else self._cloud_pos[index]

    def cloud_pos_tensor(self):
        """(N,3) float32 device tensor of the indexed positions (no host round trip)."""
        return self._pos[:self._indexed]

    def spatial_hash(self):
        return self._grid

    def _flush_input(self):
        for p, c in self._input_pending:
            self._input_pos_list.extend(p.tolist())
            self._input_rgb_list.extend(c.tolist())
        self._input_pending = []

    @property
    def _input_pos(self):
        self._flush_input()
        return self._input_pos_list

    @_input_pos.setter
    def _input_pos(self, value):           # offline tools assign a list (get_mesh_tsdf_fusion.py:67)
        self._flush_input()
        self._input_pos_list = value

    @property
    def _input_rgb(self):
        self._flush_input()
        return self._input_rgb_list

    @_input_rgb.setter
    def _input_rgb(self, value):
        self._flush_input()
        self._input_rgb_list = value

    def input_pos(self):
        return self._input_pos

    def input_rgb(self):
        return self._input_rgb

    def pts_num(self):
        return self._pts_num

    def index_train(self, xb):
        assert torch.is_tensor(xb), 'use tensor to train FAISS index'
        self.index.train(xb)
        return self.index.is_trained

    def index_ntotal(self):
        return self.index.ntotal

    def get_radius_query(self):
        return self.radius_query

    def get_geo_feats(self):
        return self.geo_feats

    def get_col_feats(self):
        return self.col_feats

    def update_geo_feats(self, feats, indices=None):
        assert torch.is_tensor(feats), 'use tensor to update features'
        if indices is not None:
            self.geo_feats[indices] = feats.detach().clone()
        else:
            assert feats.shape[0] == self.geo_feats.shape[0], 'feature shape[0] mismatch'
            self.geo_feats = feats.detach().clone()

    def update_col_feats(self, feats, indices=None):
        assert torch.is_tensor(feats), 'use tensor to update features'
        if indices is not None:
            self.col_feats[indices] = feats.detach().clone()
        else:
            assert feats.shape[0] == self.col_feats.shape[0], 'feature shape[0] mismatch'
            self.col_feats = feats.detach().clone()

    # ---- index maintenance -----------------------------------------------------------------------------------------
    def _index_add(self, pts):
        """Cover `pts` with the hash.  When they are the tail of `_pos` (the add_neural_points flow: same storage) nothing
        is copied; a foreign tensor (offline tools, get_mesh_tsdf_fusion.py:78-79) is appended to `_pos` first."""
        pts = torch.as_tensor(pts, dtype=torch.float32, device=self.device).reshape(-1, 3)
        tail = self._pos[self._indexed:]
        is_tail = pts.shape[0] == tail.shape[0] and (pts.shape[0] == 0 or pts.data_ptr() == tail.data_ptr() or
                                                     torch.equal(pts, tail))       # restore flow: a copy of the assigned list
        if not is_tail:
            assert tail.shape[0] == 0, ('index.add(x): x is neither the un-indexed tail of the cloud nor a copy of it '
                                        f'({tail.shape[0]} positions were assigned through _cloud_pos but never indexed)')
            k = pts.shape[0]
            src = pts.clone() if k else pts              # `pts` may alias the buffer that _reserve replaces
            self._pos = self._pos[:self._indexed]
            self._reserve(k, features=False)
            self._pos_buf[self._indexed:self._indexed + k] = src
            self._pos = self._pos_buf[:self._indexed + k]
            self._pos_list_cache = None
        grew = self._pos.shape[0] != self._indexed
        before = self._indexed                           # rows the previous build covered (0 after `_cloud_pos = ...`)
        self._indexed = self._pos.shape[0]
        if grew or self._grid.build_gen == 0:            # nothing appended: the hash already covers the cloud
            self._grid.build(self._pos, appended_from=before)
        self.index.is_trained = True

    def _search_all(self, q):
        """faiss-style search: D (M,8) f32, I (M,8) i64 of the 8 nearest points within the largest radius the
        configuration can ask for (slots beyond it: I=-1, D=FLT_MAX)."""
        rmax = max(self.radius_query, self.radius_add, 2.0 * self.cfg['pointcloud'].get('radius_add_max', 0.0))
        D, I, _ = ops.knn_query(self._grid, q, radius=rmax)
        return D, I.long()

    # ---- point insertion (neural_point.py:91-167) --------------------------------------------------------------------
    def reserve(self, n_points):
        """Pre-size the position / feature / hash buffers for `n_points` points, so that no append up to that size
        re-allocates (CUDA graphs captured over the cloud stay valid; see ops.SpatialHash)."""
        self._reserve(max(int(n_points) - self._pos.shape[0], 0))
        self._grid.reserve(int(n_points), self.device)

    def feature_capacity(self):
        """Rows of the feature storage behind get_geo_feats() / get_col_feats() (>= pts_num())."""
        if self._geo_buf is not None and self.geo_feats is not None and self.geo_feats.data_ptr() == self._geo_buf.data_ptr():
            return self._geo_buf.shape[0]
        return 0 if self.geo_feats is None else self.geo_feats.shape[0]

    def storage_gen(self):
        """Changes whenever a device buffer the render kernels read (positions, features, hash) was re-allocated."""
        return (self._store_gen, self._grid.alloc_gen)

    def _reserve(self, extra, features=True):
        """Room for `extra` more points behind the indexed ones in the position / feature buffers (amortised doubling).
        Tensors a caller assigned directly (`npc.geo_feats = ...`, `npc._cloud_pos = ...`) are adopted first."""
        n = self._pos.shape[0]             # indexed positions + a tail assigned through `_cloud_pos` that index.add() will cover
        need = n + int(extra)
        if (n and self._pos.data_ptr() != self._pos_buf.data_ptr()) or self._pos_buf.shape[0] < need:
            cap = max(need, 2 * self._pos_buf.shape[0], 1024)
            buf = torch.empty((cap, 3), dtype=torch.float32, device=self.device)
            buf[:n] = self._pos
            self._pos_buf, self._pos = buf, buf[:n]
            self._store_gen += 1
        if not features:
            return
        for name, bufname in (('geo_feats', '_geo_buf'), ('col_feats', '_col_buf')):
            cur, buf = getattr(self, name), getattr(self, bufname)
            rows = 0 if cur is None else cur.shape[0]
            foreign = cur is not None and (buf is None or cur.data_ptr() != buf.data_ptr())
            if buf is None or foreign or buf.shape[0] < rows + int(extra):
                cap = max(rows + int(extra), 2 * (0 if buf is None else buf.shape[0]), 1024)
                nbuf = torch.empty((cap, self.c_dim), dtype=torch.float32, device=self.device)
                if rows:
                    nbuf[:rows] = cur.detach()
                setattr(self, bufname, nbuf)
                if cur is not None:
                    setattr(self, name, nbuf[:rows])
                self._store_gen += 1

    def add_neural_points(self, batch_rays_o, batch_rays_d, batch_gt_depth, batch_gt_color, train=False,
                          is_pts_grad=False, dynamic_radius=None):
        n_rays = batch_rays_o.shape[0]
        if not n_rays:
            return 0
        self._reserve(n_rays * self.N_add)
        n0 = self._indexed
        radius = self.radius_min if is_pts_grad else self.radius_add
        counts, in_pos, in_rgb = ops.add_points(
            self._grid if self.index.is_trained else ops.SpatialHash(cell=self._grid.cell), batch_rays_o, batch_rays_d,
            batch_gt_depth, batch_gt_color, self._pos_buf[n0:], radius, dynamic_radius=dynamic_radius, n_add=self.N_add,
            fixed_interval=self.fix_interval_when_add_along_ray, near_surface=self.near_end_surface,
            far_surface=self.far_end_surface)
        n_valid, n_keep = (int(v) for v in counts.tolist())          # the one host read of this call (8 bytes)
        if dynamic_radius is not None and self.index.is_trained:
            assert n_valid == dynamic_radius.shape[0], 'shape mis-match for input points and dynamic radius'
        n_new = n_keep * self.N_add
        self._input_pending.append((in_pos[:n_keep], in_rgb[:n_keep]))
        self._pos = self._pos_buf[:n0 + n_new]
        self._pos_list_cache = None
        self._pts_num += n_new
        # fresh N(0, 0.1^2) rows drawn in the reference's order: geometry rows first, then colour rows (:150-159)
        rows = 0 if self.geo_feats is None else self.geo_feats.shape[0]
        self._geo_buf[rows:rows + n_new].normal_(mean=0, std=0.1)
        self._col_buf[rows:rows + n_new].normal_(mean=0, std=0.1)
        self.geo_feats = self._geo_buf[:rows + n_new]
        self.col_feats = self._col_buf[:rows + n_new]
        pts = self._pos[n0:]
        if n_new or not self.index.is_trained:           # nothing kept: the hash already covers the cloud (static camera)
            self.index.train(pts)
            self.index.add(pts)
        return torch.tensor(n_keep, device=self.device)

    def append_points(self, pts, geo_rows, col_rows):
        """Replica-side append: positions + their feature rows as produced by the mapping rank's add_neural_points
        (used by point_slam_b200.parallel.apply_delta; no RNG is consumed here)."""
        pts = pts.to(self.device).float().reshape(-1, 3)
        k = pts.shape[0]
        self._reserve(k)
        n0 = self._indexed
        self._pos_buf[n0:n0 + k] = pts
        self._pos = self._pos_buf[:n0 + k]
        self._pos_list_cache = None
        self._pts_num += k
        rows = 0 if self.geo_feats is None else self.geo_feats.shape[0]
        self._geo_buf[rows:rows + k] = geo_rows.to(self.device)
        self._col_buf[rows:rows + k] = col_rows.to(self.device)
        self.geo_feats, self.col_feats = self._geo_buf[:rows + k], self._col_buf[:rows + k]
        new = self._pos[n0:]
        self.index.train(new)
        self.index.add(new)

    # ---- kNN (neural_point.py:169-215) --------------------------------------------------------------------------------
    def find_neighbors_faiss(self, pos, step='add', retrain=False, is_pts_grad=False, dynamic_radius=None):
        assert step in ['add', 'query']
        if step == 'query':
            radius = self.radius_query
        else:
            radius = self.radius_min if is_pts_grad else self.radius_add
        if dynamic_radius is not None:
            assert pos.shape[0] == dynamic_radius.shape[0], 'shape mis-match for input points and dynamic radius'
        D, I, neighbor_num = ops.knn_query(self._grid, pos.float(), radius=radius, dynamic_radius=dynamic_radius, group=1)
        return D, I.long(), neighbor_num

    # ---- zero-depth rays (neural_point.py:217-277) ---------------------------------------------------------------------
    def sample_near_pcl(self, rays_o, rays_d, near, far, num):
        rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
        n_rays = rays_d.shape[0]
        intervals = 25
        far_f = far.item() if torch.is_tensor(far) else float(far)
        z_probe = torch.linspace(near, far_f, steps=intervals, device=self.device)
        pts = (rays_o[..., None, :] + rays_d[..., None, :] * z_probe[..., :, None]).reshape(-1, 3)
        _, _, neighbor_num = self.find_neighbors_faiss(pts, step='query')
        hit = neighbor_num.reshape(n_rays, intervals) > 0
        invalid = hit.sum(-1) < 2                                            # fewer than two probes near the cloud
        # first and second hit probe per ray; float64 linspace between their depths, like the numpy reference
        order = torch.argsort((~hit).to(torch.uint8), dim=1, stable=True)[:, :2]
        sec = torch.from_numpy(np.linspace(near, far_f, intervals)).to(self.device)
        z0, z1 = sec[order[:, 0]], sec[order[:, 1]]
        steps = torch.arange(num, device=self.device, dtype=torch.float64)
        z_valid = _np_linspace(z0, z1, num, steps)
        z_default = torch.from_numpy(np.linspace(near, far_f, num)).to(self.device).expand(n_rays, num)
        z = torch.where(invalid[:, None], z_default, z_valid)
        return z.float(), invalid


def _np_linspace(a, b, num, steps):
    """numpy.linspace(a, b, num) for vectors a, b in float64: a + arange(num)*step with the last sample pinned to b."""
    div = num - 1
    step = (b - a) / div
    y = a[:, None] + steps[None, :] * step[:, None]
    y[:, -1] = b
    return y
