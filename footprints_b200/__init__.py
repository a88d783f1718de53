"""footprints_b200 — B200-native (sm_100a) implementation of the Footprints encoder-decoder hot path.

Public surface mirrors the reference (nianticlabs/footprints): ``FootprintNetwork`` (alias
``FootprintsModel``), ``ModelManager`` and ``LossManager``.  The arithmetic lives in
``lib/libfootprints_b200.so`` (C ABI: ``include/footprints_b200.h``).
"""
from .network import FootprintNetwork, FootprintsModel, ResnetEncoder, SkipDecoder  # noqa: F401
from .model_manager import ModelManager  # noqa: F401
from .losses import LossManager  # noqa: F401
from .segmentor import Segmentor  # noqa: F401

__all__ = ['FootprintNetwork', 'FootprintsModel', 'ResnetEncoder', 'SkipDecoder', 'ModelManager', 'LossManager', 'Segmentor']
