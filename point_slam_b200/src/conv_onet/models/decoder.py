"""Drop-in for the reference `src/conv_onet/models/decoder.py`: same class names, constructor signatures,
sub-module / parameter names (state_dict keys) and parameter creation ORDER (so a seeded construction gives the
same initial weights as the reference), but `forward` runs the fused sm_100a kernels of libpointslam_b200.so.

Reference lines each piece stands in for are cited inline (`decoder.py:N` = reference file).
"""
import math

import torch
import torch.nn as nn

from ....ops import RenderSettings, decode, decoder_param_list
from .... import _lib as L


class GaussianFourierFeatureTransform(nn.Module):
    """Parameter holder for a random Fourier basis B (decoder.py:8-37).  Learnable -> nn.Parameter, otherwise a
    plain tensor attribute that is NOT in the state_dict (decoder.py:27-28), exactly like the reference."""

    def __init__(self, num_input_channels, mapping_size=93, scale=25, learnable=False, concat=True):
        super().__init__()
        self.concat, self.mapping_size, self.scale, self.learnable = concat, mapping_size, scale, learnable
        basis = torch.randn((num_input_channels, mapping_size)) * scale
        self._B = nn.Parameter(basis) if learnable else basis

    def _apply(self, fn, *a, **kw):                 # keep the non-registered basis on the module's device
        out = super()._apply(fn, *a, **kw)
        if not self.learnable:
            self._B = fn(self._B)
        return out

    def forward(self, x):                           # stand-alone use only (the render path embeds inside the kernel)
        x = x.squeeze(0)
        y = (2 * math.pi * x) @ self._B.to(x.device)
        return torch.cat((torch.sin(y), torch.cos(y)), dim=-1) if self.concat else torch.sin(y)


class DenseLayer(nn.Linear):
    """nn.Linear with Xavier-uniform weight (gain by activation name) and zero bias (decoder.py:40-49)."""

    def __init__(self, in_dim, out_dim, activation='relu', *args, **kwargs):
        self.activation = activation
        super().__init__(in_dim, out_dim, *args, **kwargs)

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight, gain=nn.init.calculate_gain(self.activation))
        if self.bias is not None:
            nn.init.zeros_(self.bias)


class MLP_col_neighbor(nn.Module):
    """F_theta: (c_dim + 2*10) -> hidden -> c_dim, Softplus(beta=100) (decoder.py:225-240)."""

    def __init__(self, c_dim, embedding_size_rel, hidden_size):
        super().__init__()
        self.linear1 = nn.Linear(c_dim + embedding_size_rel, hidden_size)
        self.linear2 = nn.Linear(hidden_size, c_dim)
        self.act_fn = nn.Softplus(beta=100)
        nn.init.xavier_uniform_(self.linear1.weight)
        nn.init.xavier_uniform_(self.linear2.weight)

    def forward(self, x):
        return self.linear2(self.act_fn(self.linear1(x)))


class MLP_exposure(nn.Module):
    """Exposure latent (8) -> 12 affine coefficients (decoder.py:243-258).  Eight inputs: evaluated with torch."""

    def __init__(self, latent_dim, hidden_size):
        super().__init__()
        self.linear1 = nn.Linear(latent_dim, hidden_size)
        self.linear2 = nn.Linear(hidden_size, 12)
        self.act_fn = nn.Softplus(beta=100)
        nn.init.normal_(self.linear1.weight, mean=0, std=0.01)
        nn.init.normal_(self.linear2.weight, mean=0, std=0.01)

    def forward(self, x):
        return self.linear2(self.act_fn(self.linear1(x)))


def _trunk(embedding_input, hidden_size, n_blocks, skips):
    layers = [DenseLayer(embedding_input, hidden_size, activation='relu')]
    for i in range(n_blocks - 1):
        layers.append(DenseLayer(hidden_size + (embedding_input if i in skips else 0), hidden_size, activation='relu'))
    return nn.ModuleList(layers)


class MLP_geometry(nn.Module):
    """Geometry decoder parameters (decoder.py:62-128): 93-d sin embedding, 5 x 32 ReLU trunk with per-layer
    feature injection, skip-cat after block 2, 1-d occupancy logit."""

    def __init__(self, cfg, c_dim=32, hidden_size=128, n_blocks=5, leaky=False, sample_mode='bilinear', skips=[2],
                 pos_embedding_method='fourier', concat_feature=False, use_view_direction=False):
        super().__init__()
        assert pos_embedding_method == 'fourier' and not use_view_direction and n_blocks == 5 and list(skips) == [2]
        self.feat_name = 'geometry_feat'
        self.c_dim, self.n_blocks, self.skips = c_dim, n_blocks, skips
        self.weighting = cfg['pointcloud']['nn_weighting']
        self.use_dynamic_radius = cfg['use_dynamic_radius']
        self.min_nn_num = cfg['pointcloud']['min_nn_num']
        self.N_surface = cfg['rendering']['N_surface']
        self.use_view_direction = use_view_direction
        # creation order below == reference (RNG parity of seeded construction)
        self.fc_c = nn.ModuleList([nn.Linear(c_dim, hidden_size) for _ in range(n_blocks)])
        self.embedder = GaussianFourierFeatureTransform(3, mapping_size=93, scale=25, concat=False, learnable=True)
        self.embedder_rel_pos = GaussianFourierFeatureTransform(3, mapping_size=10, scale=32, learnable=True)
        self.mlp_col_neighbor = MLP_col_neighbor(c_dim, 2 * self.embedder_rel_pos.mapping_size, hidden_size)  # unused, kept for state_dict parity
        self.pts_linears = _trunk(93, hidden_size, n_blocks, skips)
        self.output_linear = DenseLayer(hidden_size, 1, activation='relu')
        self.actvn = nn.Softplus(beta=100)          # present but unused by the reference trunk too (decoder.py:211)
        self.sample_mode = sample_mode


class MLP_color(nn.Module):
    """Colour decoder parameters (decoder.py:261-339): [sin,cos] 40-d embedding (non-learnable B), optional
    per-neighbour MLP on rel-pos encoded features, 5 x 128 Softplus(100) trunk, 3-d output."""

    def __init__(self, cfg, c_dim=32, hidden_size=128, n_blocks=5, leaky=False, sample_mode='bilinear', skips=[2],
                 pos_embedding_method='fourier', concat_feature=False, use_view_direction=False):
        super().__init__()
        assert pos_embedding_method == 'fourier' and n_blocks == 5 and list(skips) == [2]
        if use_view_direction:
            raise NotImplementedError('use_view_direction=True is off in every shipped config (point_slam.yaml:15) '
                                      'and is not on the B200 hot path yet')
        self.feat_name = 'color_feat'
        self.c_dim, self.n_blocks, self.skips = c_dim, n_blocks, skips
        self.weighting = cfg['pointcloud']['nn_weighting']
        self.min_nn_num = cfg['pointcloud']['min_nn_num']
        self.use_dynamic_radius = cfg['use_dynamic_radius']
        self.N_surface = cfg['rendering']['N_surface']
        self.use_view_direction = use_view_direction
        self.encode_rel_pos_in_col = cfg['model']['encode_rel_pos_in_col']
        self.encode_exposure = cfg['model']['encode_exposure']
        self.encode_viewd = cfg['model']['encode_viewd']
        self.fc_c = nn.ModuleList([nn.Linear(c_dim, hidden_size) for _ in range(n_blocks)])
        self.embedder = GaussianFourierFeatureTransform(3, mapping_size=20, scale=32)
        self.embedder_rel_pos = GaussianFourierFeatureTransform(3, mapping_size=10, scale=32, learnable=True)
        self.mlp_col_neighbor = MLP_col_neighbor(c_dim, 2 * self.embedder_rel_pos.mapping_size, hidden_size)
        if self.encode_exposure:
            self.mlp_exposure = MLP_exposure(cfg['model']['exposure_dim'], hidden_size)
        self.pts_linears = _trunk(40, hidden_size, n_blocks, skips)
        self.output_linear = DenseLayer(hidden_size, 3, activation='linear')
        self.actvn = nn.Softplus(beta=100)
        self.sample_mode = sample_mode


class POINT(nn.Module):
    """Decoder for point-represented features (decoder.py:452-518).  `forward` keeps the reference's positional
    signature (it is called positionally from Renderer.eval_points, Renderer.py:60-62)."""

    def __init__(self, cfg, c_dim=32, hidden_size=128, pos_embedding_method='fourier', use_view_direction=False):
        super().__init__()
        assert c_dim == 32 and hidden_size == 128, 'kernels are specialised for c_dim=32, hidden 32/128'
        self.geo_decoder = MLP_geometry(cfg=cfg, c_dim=c_dim, skips=[2], n_blocks=5, hidden_size=32,
                                        pos_embedding_method=pos_embedding_method)
        self.color_decoder = MLP_color(cfg=cfg, c_dim=c_dim, skips=[2], n_blocks=5, hidden_size=hidden_size,
                                       pos_embedding_method=pos_embedding_method, use_view_direction=use_view_direction)
        self._radius_query = cfg['pointcloud']['radius_query']

    # -- helpers shared with the fused Renderer path ------------------------------------------------------------
    def kernel_params(self):
        return decoder_param_list(self)

    def settings(self, stage, S, is_tracker, coef=0.1, near_surface=0.98, far_surface=1.02, exposure_feat=None,
                 radius_query=None):
        cd = self.color_decoder
        mode = L.RGB_SIGMOID
        if cd.encode_exposure:
            mode = L.RGB_AFFINE_SIGMOID if exposure_feat is not None else L.RGB_RAW
        return RenderSettings(stage=stage, S=S, near_surface=near_surface, far_surface=far_surface,
                              radius_query=self._radius_query if radius_query is None else radius_query, coef=coef,
                              encode_rel_pos=cd.encode_rel_pos_in_col, rgb_mode=mode, weighting=cd.weighting,
                              min_nn=cd.min_nn_num, is_tracker=is_tracker)

    def exposure_affine(self, exposure_feat):
        """12 affine coefficients from the exposure latent (decoder.py:433-436); None when not applicable."""
        if self.color_decoder.encode_exposure and exposure_feat is not None:
            return self.color_decoder.mlp_exposure(exposure_feat).reshape(-1)
        return None

    def draw_no_neighbor_vectors(self, stage, device):
        """The N(0, 0.01^2) vectors given to samples without neighbours, drawn from the device RNG in the reference's
        order: geometry (decoder.py:170-171) then, in the colour stage, colour (:387-388)."""
        rg = torch.zeros([32], device=device).normal_(mean=0, std=0.01)
        rc = torch.zeros([32], device=device).normal_(mean=0, std=0.01) if stage == 'color' else None
        return rg, rc

    def forward(self, p, npc, stage, npc_geo_feats, npc_col_feats, pts_num=16, is_tracker=False, cloud_pos=None,
                pts_views_d=None, dynamic_r_query=None, exposure_feat=None):
        assert stage in ('geometry', 'color')
        pts = p.reshape(-1, 3)
        dev = pts.device
        if cloud_pos is None:
            cloud_pos = npc.cloud_pos_tensor()
        st = self.settings(stage, pts_num, is_tracker, exposure_feat=exposure_feat, radius_query=npc.get_radius_query())
        r2 = None
        if self.geo_decoder.use_dynamic_radius:
            r2 = (dynamic_r_query.detach().reshape(-1).to(device=dev, dtype=torch.float64) ** 2).contiguous()
        rg, rc = self.draw_no_neighbor_vectors(stage, dev)
        raw, point_mask = decode(st, npc.spatial_hash(), self.kernel_params(), pts, cloud_pos, npc_geo_feats,
                                 npc_col_feats if stage == 'color' else None, r2_pts=r2, r2_group=1, rand_geo=rg,
                                 rand_col=rc, affine=self.exposure_affine(exposure_feat))
        # a ray is valid if at least int(N_surface/2+1) of its samples have neighbours (decoder.py:200-201)
        ray_mask = point_mask.view(-1, pts_num).sum(1) >= int(self.geo_decoder.N_surface / 2 + 1)
        return raw, ray_mask, point_mask
