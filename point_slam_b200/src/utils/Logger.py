"""Drop-in for the reference `src/utils/Logger.py` (Logger.py:6-44): same class, constructor, `log(...)` signature, file name
`ckpts/{idx:05d}.tar` and the same 14 dictionary keys with the same types, so that `src/tools/eval_ate.py`,
`get_mesh_tsdf_fusion.py` and a resumed run keep reading the files.  Differences, both invisible to readers of the file:

* `geo_feats` / `col_feats` are written as exact-size tensors (the drop-in NeuralPointCloud hands out views of its
  capacity-doubling buffers; `torch.save` of a view would serialise the whole buffer);
* `cloud_pos` stays a Python list of [x, y, z] like the reference writes (get_mesh_tsdf_fusion.py:66,77 rebuilds a tensor
  from it); it is produced by one `tolist()` of the device tensor instead of being maintained as a list.

`load_neural_point_cloud` is the restore step of get_mesh_tsdf_fusion.py:64-82 for the drop-in class.
"""
import os

import torch


# key order of the dictionary the reference writes (Logger.py:22-41)
CKPT_KEYS = ('geo_feats', 'col_feats', 'cloud_pos', 'pts_num', 'input_pos', 'input_rgb', 'decoder_state_dict', 'gt_c2w_list',
             'estimate_c2w_list', 'keyframe_list', 'keyframe_dict', 'selected_keyframes', 'idx', 'exposure_feat_all')


def checkpoint_dict(idx, npc, decoders, gt_c2w_list, estimate_c2w_list, keyframe_dict, keyframe_list, selected_keyframes,
                    exposure_feat=None):
    """The checkpoint payload: CKPT_KEYS -> values with the reference's types (tensors, Python lists, ints, state dict)."""
    exposure = None if exposure_feat is None else torch.stack(exposure_feat, dim=0)
    values = (npc.get_geo_feats().detach().clone(), npc.get_col_feats().detach().clone(), npc.cloud_pos(), npc.pts_num(),
              npc.input_pos(), npc.input_rgb(), decoders.state_dict(), gt_c2w_list, estimate_c2w_list, keyframe_list,
              keyframe_dict, selected_keyframes, idx, exposure)
    return dict(zip(CKPT_KEYS, values))


class Logger(object):
    """Checkpoint writer with the reference's constructor and `log` signature (called from Mapper.py:775-777)."""

    def __init__(self, cfg, args, mapper):
        self.verbose, self.ckptsdir = mapper.verbose, mapper.ckptsdir
        self.gt_c2w_list, self.estimate_c2w_list = mapper.gt_c2w_list, mapper.estimate_c2w_list
        self.decoders = mapper.decoders

    def log(self, idx, keyframe_dict, keyframe_list, selected_keyframes, npc, exposure_feat=None):
        path = os.path.join(self.ckptsdir, f'{idx:05d}.tar')
        torch.save(checkpoint_dict(idx, npc, self.decoders, self.gt_c2w_list, self.estimate_c2w_list, keyframe_dict,
                                   keyframe_list, selected_keyframes, exposure_feat), path)
        if self.verbose:
            print('Saved checkpoints at', path)
        return path


def load_neural_point_cloud(npc, ckpt, device):
    """Restore a NeuralPointCloud from a checkpoint dictionary written by `Logger.log` (this one or the reference's):
    the assignments of get_mesh_tsdf_fusion.py:66-80 followed by the index rebuild.  Returns the number of indexed points."""
    npc._cloud_pos = ckpt['cloud_pos']
    npc._input_pos = ckpt['input_pos']
    npc._input_rgb = ckpt['input_rgb']
    npc._pts_num = len(ckpt['cloud_pos'])
    npc.geo_feats = ckpt['geo_feats'].to(device)
    npc.col_feats = ckpt['col_feats'].to(device)
    cloud_pos = torch.tensor(ckpt['cloud_pos'], device=device)
    npc.index_train(cloud_pos)
    npc.index.add(cloud_pos)
    assert npc.geo_feats.shape[0] == npc.col_feats.shape[0] == npc.index_ntotal() == npc._pts_num, \
        'checkpoint is inconsistent: feature rows / positions / pts_num differ'
    return npc.index_ntotal()
