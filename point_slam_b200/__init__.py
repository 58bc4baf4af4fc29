"""point_slam_b200 -- B200-native (sm_100a) implementation of Point-SLAM's per-frame volumetric rendering hot path.

`point_slam_b200.src.*` mirrors the reference's module paths for the path (`src.neural_point`,
`src.conv_onet.models.decoder`, `src.utils.Renderer`, `src.common`); everything there calls hand-written CUDA
kernels through the C ABI in include/pointslam_b200.h (`point_slam_b200/libpointslam_b200.so`).
"""
__version__ = '0.1.0'
