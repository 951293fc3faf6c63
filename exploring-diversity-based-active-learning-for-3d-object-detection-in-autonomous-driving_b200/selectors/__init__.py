"""Selector package with the reference's public names (det3d/selectors/__init__.py).

Only the data-only diversity selectors are provided; the detector-driven ones (feature, entropy,
badge, ...) need a trained model and are out of this hot path's scope.
"""
from .base_selector import BaseSelector
from .bev_spatial_selector import BevSpatialSelector
from .builder import build_selector
from .euclidean_spatial_selector import EuSpatialSelector
from .registry import SELECTORS
from .spatial_selector import SpatialSelector
from .spatial_temporal_selector import SpatialTemporalSelector
from .temporal_selector import TemporalSelector

__all__ = ["BaseSelector", "SpatialSelector", "EuSpatialSelector", "TemporalSelector",
           "SpatialTemporalSelector", "BevSpatialSelector", "SELECTORS", "build_selector"]
