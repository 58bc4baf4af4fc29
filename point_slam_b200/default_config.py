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
    # caller-side keys the iteration shells read (Tracker.py:46-60, Mapper.py:40-80); defaults of configs/point_slam.yaml:23-83
    'mapping': {'device': 'cuda:0', 'BA': False, 'geo_iter_ratio': 0.4, 'geo_iter_first': 400, 'every_frame': 5, 'pixels': 1000,
                'iters': 400, 'iters_first': 1500, 'pixels_adding': 6000, 'pixels_based_on_color_grad': 0, 'w_color_loss': 0.1,
                'frustum_edge': -4, 'min_iter_ratio': 0.95, 'mapping_window_size': 5, 'fix_geo_decoder': True,
                'fix_color_decoder': False, 'frustum_feature_selection': True, 'keyframe_every': 50,
                'init': {'geometry': {'decoders_lr': 0.001, 'geometry_lr': 0.03, 'color_lr': 0.0},
                         'color': {'decoders_lr': 0.005, 'geometry_lr': 0.005, 'color_lr': 0.005}},
                'stage': {'geometry': {'decoders_lr': 0.001, 'geometry_lr': 0.03, 'color_lr': 0.0},
                          'color': {'decoders_lr': 0.005, 'geometry_lr': 0.005, 'color_lr': 0.005}}},
    'tracking': {'device': 'cuda:0', 'lr': 0.002, 'pixels': 200, 'iters': 20, 'w_color_loss': 0.5, 'separate_LR': True,
                 'handle_dynamic': True, 'use_color_in_tracking': True, 'depth_limit': False, 'sample_with_color_grad': False,
                 'ignore_edge_W': 20, 'ignore_edge_H': 20, 'const_speed_assumption': True, 'gt_camera': False},
    'cam': {'H': 480, 'W': 640, 'fx': 517.3, 'fy': 516.5, 'cx': 318.6, 'cy': 255.3, 'crop_edge': 0},
    'rendering': {'N_surface': 5, 'near_end': 0.3, 'near_end_surface': 0.98, 'far_end_surface': 1.02,
                  'sigmoid_coef_tracker': 0.1, 'sigmoid_coef_mapper': 0.1, 'sample_near_pcl': True},
    'pointcloud': {'nn_num': 8, 'min_nn_num': 2, 'N_add': 3, 'nn_weighting': 'distance', 'radius_add': 0.04,
                   'radius_min': 0.02, 'radius_query': 0.08, 'radius_add_max': 0.08, 'radius_add_min': 0.02,
                   'radius_query_ratio': 2, 'color_grad_threshold': 0.15, 'near_end_surface': 0.98,
                   'far_end_surface': 1.02, 'nlist': 400, 'nprobe': 4, 'fix_interval_when_add_along_ray': False},
}
_DATASET = {
    'replica': {'rendering': {'sample_near_pcl': False},
                'tracking': {'ignore_edge_W': 100, 'ignore_edge_H': 100, 'pixels': 1500, 'iters': 40},
                'mapping': {'keyframe_every': 20, 'mapping_window_size': 12, 'pixels': 5000, 'pixels_based_on_color_grad': 1000,
                            'iters': 300}},
    'tum': {'model': {'encode_rel_pos_in_col': False},
            'tracking': {'separate_LR': False, 'pixels': 5000, 'iters': 200, 'sample_with_color_grad': True},
            'mapping': {'every_frame': 2, 'mapping_window_size': 10, 'pixels': 10000, 'iters_first': 500, 'geo_iter_first': 200,
                        'iters': 150}},
    'scannet': {'model': {'encode_exposure': True, 'encode_rel_pos_in_col': False, 'encode_viewd': False},
                'rendering': {'near_end_surface': 0.96, 'far_end_surface': 1.04},
                'pointcloud': {'near_end_surface': 0.96, 'far_end_surface': 1.04},
                'tracking': {'separate_LR': False, 'lr': 0.0005, 'pixels': 5000, 'iters': 100, 'sample_with_color_grad': True},
                'mapping': {'geo_iter_ratio': 0.3, 'mapping_window_size': 20, 'keyframe_every': 10, 'pixels': 10000,
                            'iters_first': 500, 'geo_iter_first': 200, 'iters': 300}},
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
