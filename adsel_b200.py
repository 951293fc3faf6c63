"""Import alias for the product package.

The package directory is named after the reference repository
(``exploring-diversity-based-active-learning-for-3d-object-detection-in-autonomous-driving_b200``),
which is not a valid Python identifier.  ``import adsel_b200`` loads that directory as a regular
package under the importable name ``adsel_b200`` (so ``from adsel_b200.selectors import
build_selector`` works exactly like ``from det3d.selectors import build_selector`` does in the
reference).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(
    os.path.dirname(os.path.abspath(__file__)),
    "exploring-diversity-based-active-learning-for-3d-object-detection-in-autonomous-driving_b200",
)


def _load():
    spec = importlib.util.spec_from_file_location(
        "adsel_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
    )
    module = importlib.util.module_from_spec(spec)
    sys.modules["adsel_b200"] = module  # replaces this shim while the package initialises
    spec.loader.exec_module(module)
    return module


_load()
