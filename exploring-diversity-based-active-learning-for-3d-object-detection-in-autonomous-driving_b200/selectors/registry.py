"""``SELECTORS`` registry (det3d/selectors/registry.py:3)."""
from ..utils import Registry

SELECTORS = Registry("selector")
