"""pcgol_b200 — B200-native (sm_100a) drop-in for the data-parallel hot path of
seqsense/pcgol: storage.Search (Nearest/Range), filter.VoxelGrid and the
point-to-point ICP loop.  Host-side mirror of the reference's interfaces over the
C ABI in include/pcgol_b200.h; all compute is in libpcgol_b200.so (hand-written CUDA).
"""
from . import _lib
from ._lib import PcgError, device_count, kernel_launch_count, E_INVALID_ARG
from .pc import PointCloud, PointCloudHeader
from .storage import Index, Neighbor
from .filter import VoxelGrid, NoPointError, ReferencePanic
from .segmentation import RegionGrowing
from . import io
from .io import DeviceCloud
from .icp import (ErrNotEnoughPairs, Evaluated, GaussNewtonUpdaterFactory, GradientDescentUpdaterFactory,
                  NearestPointCorresponder, PointToPointEvaluator, PointToPointICPGradient, Stat, STRICT, FAST,
                  WITH_HESSIAN)

__all__ = [
    "PcgError", "device_count", "kernel_launch_count", "PointCloud", "PointCloudHeader", "Index", "Neighbor",
    "VoxelGrid", "NoPointError", "ReferencePanic", "ErrNotEnoughPairs", "Evaluated",
    "GradientDescentUpdaterFactory", "NearestPointCorresponder", "PointToPointEvaluator",
    "PointToPointICPGradient", "Stat", "STRICT", "FAST", "WITH_HESSIAN", "GaussNewtonUpdaterFactory", "E_INVALID_ARG",
    "RegionGrowing", "DeviceCloud", "io",
]
