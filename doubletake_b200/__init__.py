"""doubletake_b200: B200-native (sm_100a) plane-sweep MVS depth engine -- a drop-in for the hot path of
nianticlabs/doubletake (CostVolumeManager family -> CVEncoder -> depth decoder, behind DepthModel*.forward).

Host code is Python/PyTorch plumbing; every arithmetic op on the path runs in hand-written CUDA reached through the
C ABI of include/doubletake_b200.h (doubletake_b200/libdoubletake_b200.so).  There is no CPU / PyTorch fallback.
"""
from .cost_volume import (CostVolumeManager, FastFeatureMeshHintVolumeManager, FeatureMeshHintVolumeManager,  # noqa: F401
                          FeatureVolumeManager, MLP, to_b200)
from .depth_model import DepthModel, DepthModelCVHint, HotPathOptions, install  # noqa: F401
from .networks import BasicBlock, ConvPlan, CVEncoder, DepthDecoderPP, SkipDecoderRegression  # noqa: F401
from .encoders import ResnetMatchingEncoder  # noqa: F401
from . import formats  # noqa: F401
from .tsdf import TSDF, TSDFFuser, get_frustum_bounds  # noqa: F401

__all__ = [
    "CostVolumeManager", "FeatureVolumeManager", "FeatureMeshHintVolumeManager", "FastFeatureMeshHintVolumeManager",
    "MLP", "to_b200", "DepthModel", "DepthModelCVHint", "HotPathOptions", "install", "BasicBlock", "ConvPlan",
    "CVEncoder", "DepthDecoderPP", "SkipDecoderRegression", "TSDF", "TSDFFuser", "get_frustum_bounds", "ResnetMatchingEncoder",
    "formats",
]
