"""reference: configuration/mmsdnet_config_chaos.py (same keys and values)"""
from ..loaders.synthetic_chaos import SyntheticChaosLoader

params = {
    'seed': 10,
    'folder': 'mmsdnet_chaos',
    'epochs': 500,
    'batch_size': 6,
    'split': 0,
    'dataset_name': 'chaos',
    'test_dataset': 'chaos',
    'image_downsample': 1,
    'modality': ['t1', 't2'],
    'model': 'mmsdnet.MMSDNet',
    'executor': 'mmsdnet_executor.MMSDNetExecutor',
    'l_mix': 1,
    'decoder_type': 'film',
    'num_z': 8,
    'w_sup_M': 10,
    'w_adv_M': 1,
    'w_rec_X': 10,
    'w_adv_X': 1,
    'w_rec_Z': 1,
    'w_kl': 0.1,
    'lr': 0.0001,
}

d_mask_params = {'filters': 4, 'lr': 0.0001, 'name': 'D_Mask'}

anatomy_encoder_params = {
    'normalise': 'batch',
    'downsample': 4,
    'filters': 64,
    'out_channels': 8,
    'rounding': True,
}


def get(input_shape=None):
    p = dict(params)
    dm, ae = dict(d_mask_params), dict(anatomy_encoder_params)
    loader = SyntheticChaosLoader()
    shp = tuple(input_shape) if input_shape is not None else tuple(loader.input_shape)
    ratio = p['image_downsample']
    shp = (int(shp[0] / ratio), int(shp[1] / ratio), shp[2])
    p['input_shape'] = shp
    p['num_masks'] = loader.num_masks
    dm['input_shape'] = (shp[:-1]) + (loader.num_masks,)
    ae['input_shape'] = shp
    ae['output_shape'] = (shp[:-1]) + (ae['out_channels'],)
    p.update({'anatomy_encoder': ae, 'd_mask_params': dm})
    return p
