"""fdbm_b200 -- B200-native enhancement hot path behind the reference's Python API.

    from fdbm_b200 import BackboneRegistry, Bridge, SpecsDataModule, EnhancementModel, pad_spec

Everything tensor-sized runs in libfdbm_b200.so (hand-written sm_100a CUDA, see ../csrc and
include/fdbm_b200.h); this package is the host-side mirror of the reference's `fdbm` interfaces.
"""
from .registry import BackboneRegistry, BridgeRegistry, Registry
from .bridge import Bridge, ProbabilityPathFM, ProbabilityPathSB
from .data_module import SpecsDataModule, get_window, pad_spec, padded_frames
from .backbones import NCSNpp_v2, NCSNpp_v2_predictive, sensitise_
from .tfgridnet import TFGridNet_4l32c80, TFGridNet_5l32c100, TFGridNet_5l32c100_predictive
from .model import (EnhancementModel, PredictiveEnhancementModel, si_sdr, split_list, shard_for_rank, gather_waveforms,
                    length_buckets)

__all__ = [
    "BackboneRegistry", "BridgeRegistry", "Registry", "Bridge", "ProbabilityPathSB", "ProbabilityPathFM",
    "SpecsDataModule", "get_window", "pad_spec", "padded_frames", "NCSNpp_v2", "NCSNpp_v2_predictive", "sensitise_",
    "TFGridNet_5l32c100", "TFGridNet_4l32c80", "TFGridNet_5l32c100_predictive",
    "EnhancementModel", "PredictiveEnhancementModel", "si_sdr", "split_list", "shard_for_rank", "gather_waveforms",
    "length_buckets",
]
