"""The configuration keys the hot path reads ("Config / flags"), with the reference's default
values (configs/point_slam.yaml) and the per-dataset overrides that change kernel shape or branches
(configs/Replica/replica.yaml, configs/TUM_RGBD/tum.yaml, configs/ScanNet/scannet.yaml).  The reference YAML files
themselves are reused unchanged by a full SLAM run; this dict exists so tests and bench.py need no YAML tree."""
import copy

_BASE = {
    'setup_seed': 1219,
    'use_dynamic_radius': True,
    'model': {'c_dim': 32, 'exposure_dim': 8, 'pos_embedding_method': 'fourier', 'encode_rel_pos_in_col': True,
              'encode_exposure': False, 'use_view_direction': False, 'encode_viewd': True},
    'mapping': {'device': 'cuda:0'},
    'tracking': {'device': 'cuda:0'},
    'cam': {'H': 480, 'W': 640, 'fx': 517.3, 'fy': 516.5, 'cx': 318.6, 'cy': 255.3, 'crop_edge': 0},
    'rendering': {'N_surface': 5, 'near_end': 0.3, 'near_end_surface': 0.98, 'far_end_surface': 1.02,
                  'sigmoid_coef_tracker': 0.1, 'sigmoid_coef_mapper': 0.1, 'sample_near_pcl': True},
    'pointcloud': {'nn_num': 8, 'min_nn_num': 2, 'N_add': 3, 'nn_weighting': 'distance', 'radius_add': 0.04,
                   'radius_min': 0.02, 'radius_query': 0.08, 'radius_add_max': 0.08, 'radius_add_min': 0.02,
                   'radius_query_ratio': 2, 'color_grad_threshold': 0.15, 'near_end_surface': 0.98,
                   'far_end_surface': 1.02, 'nlist': 400, 'nprobe': 4, 'fix_interval_when_add_along_ray': False},
}
_DATASET = {
    'replica': {'rendering': {'sample_near_pcl': False}},
    'tum': {'model': {'encode_rel_pos_in_col': False}},
    'scannet': {'model': {'encode_exposure': True, 'encode_rel_pos_in_col': False, 'encode_viewd': False},
                'rendering': {'near_end_surface': 0.96, 'far_end_surface': 1.04},
                'pointcloud': {'near_end_surface': 0.96, 'far_end_surface': 1.04}},
}


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v


def make_cfg(dataset='replica', device='cuda:0', **overrides):
    """overrides use dotted keys, e.g. make_cfg('replica', **{'rendering.N_surface': 32})."""
    cfg = copy.deepcopy(_BASE)
    _merge(cfg, _DATASET[dataset])
    cfg['mapping']['device'] = cfg['tracking']['device'] = device
    for path, val in overrides.items():
        node = cfg
        keys = path.split('.')
        for k in keys[:-1]:
            node = node.setdefault(k, {})
        node[keys[-1]] = val
    return cfg
