"""B200-native physics-inspired ultrasound augmentation (hot path of
mariatirindelli/UltrasoundAugmentation, MICCAI 2021 "Rethinking Ultrasound Augmentation").

Host code is Python (transform callables + parameter sampling); all compute runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/usaug.h (libusaug.so).
There is no CPU fallback and no other backend.
"""
from . import geometry, ops, params, sharding, synth  # noqa: F401
from .geometry import ConvexGeometry
from .transforms import (ConfidenceSNR, GraphedChain, PhysicsChain, ProbePressureDeformation, Reverberation,
                         TimeGainCompensation)

__all__ = ["ProbePressureDeformation", "TimeGainCompensation", "ConfidenceSNR", "Reverberation",
           "PhysicsChain", "GraphedChain", "ConvexGeometry", "geometry", "ops", "params", "sharding", "synth"]
__version__ = "1.2.0"
