"""Factories with the reference's signatures (lib/memory/build.py:5-32).

Extra, optional config keys (absent keys keep upstream behaviour):
  cfg.CONTRAST.QUEUE_DTYPE  'fp32' (default, parity mode) | 'bf16' (tcgen05 path)
  cfg.CONTRAST.ALGO         'auto' | 'ffma' | 'tcgen05'
"""
from .moco_queue import RGBMoCo, CMCMoCo
from .mem_bank import RGBMem, CMCMem
from .losses import NCESoftmaxLoss, NCECriterion, D


def _opt(node, name, default):
    try:
        return getattr(node, name)
    except (AttributeError, KeyError):
        return default


def create_contrast(cfg, n_data):
    if cfg.CONTRAST.MEM_TYPE == 'moco':
        mem_func = RGBMoCo if cfg.CROSS.MODALITY == 'visual' else CMCMoCo
        memory = mem_func(cfg.CROSS.FEAT_DIM, cfg.CONTRAST.NCE_K, cfg.CONTRAST.NCE_T,
                          queue_dtype=_opt(cfg.CONTRAST, 'QUEUE_DTYPE', 'fp32'),
                          algo=_opt(cfg.CONTRAST, 'ALGO', 'auto'))
    elif cfg.CONTRAST.MEM_TYPE == 'simsiam':
        memory = None
    elif cfg.CONTRAST.MEM_TYPE == 'bank':
        mem_func = RGBMem if cfg.CROSS.MODALITY == 'visual' else CMCMem
        memory = mem_func(cfg.CROSS.FEAT_DIM, n_data, cfg.CONTRAST.NCE_K, cfg.CONTRAST.NCE_T, cfg.CONTRAST.NCE_M)
    else:
        raise NotImplementedError('mem not suported: {}'.format(cfg.CONTRAST.MEM_TYPE))
    return memory


def create_criterion(cfg, n_data):
    if cfg.CROSS.CRITERION == 'crossentropy':
        criterion = NCESoftmaxLoss()
    elif cfg.CROSS.CRITERION == 'simsiam_d':
        criterion = D()
    elif cfg.CROSS.CRITERION == 'NCE':
        criterion = NCECriterion(n_data)
    else:
        raise NotImplementedError('criterion not suported: {}'.format(cfg.CROSS.CRITERION))
    return criterion
