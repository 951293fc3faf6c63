"""Selector package with the reference's public names (det3d/selectors/__init__.py).

The data-only diversity selectors and the random baseline, plus the feature-distance selectors (feature, spatial+feature, spatial+temporal+feature) on STORED features (``pred=False``); running
the detector to produce features / entropies (feature with ``pred=True``, entropy, badge, ...) needs a trained
model and is out of this hot path's scope.
"""
from .base_selector import BaseSelector
from .bev_spatial_selector import BevSpatialSelector
from .builder import build_selector
from .euclidean_spatial_selector import EuSpatialSelector
from .feature_selector import FeatureSelector
from .random_selector import RandomSelector
from .registry import SELECTORS
from .spatial_feature_selector import SpatialFeatureSelector, SpatialTemporalFeatureSelector
from .spatial_selector import SpatialSelector
from .spatial_temporal_selector import SpatialTemporalSelector
from .temporal_selector import TemporalSelector

__all__ = ["BaseSelector", "RandomSelector", "SpatialSelector", "EuSpatialSelector", "TemporalSelector",
           "SpatialTemporalSelector", "BevSpatialSelector", "FeatureSelector", "SpatialFeatureSelector",
           "SpatialTemporalFeatureSelector", "SELECTORS", "build_selector"]
