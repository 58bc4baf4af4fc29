"""Model factory with the reference's signature (src/conv_onet/config.py:4-21)."""
from . import models


def get_model(cfg):
    c_dim = cfg['model']['c_dim']
    decoder = models.decoder_dict['point'](cfg=cfg, c_dim=c_dim, pos_embedding_method=cfg['model']['pos_embedding_method'],
                                          use_view_direction=cfg['model']['use_view_direction'])
    return decoder
