"""oracle/reference_loader.py -- TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference selector modules from /root/reference by path:
the package ``det3d.selectors`` cannot be imported as a whole (its __init__ pulls in the detector
stack: pycocotools, spconv, compiled ops), so

1. stub modules ``addict`` (a dict with attribute access, named ``Dict``) and ``terminaltables``
   (an empty ``AsciiTable``) are registered -- both are absent from this image and only touched by
   ``det3d.torchie`` at import time;
2. an empty ``det3d.selectors`` package object with the real ``__path__`` is registered so the
   package ``__init__`` is skipped;
3. the selector modules are then imported normally, e.g.
   ``det3d.selectors.spatial_temporal_selector``.

Where /root/reference does not exist (the GPU box) the same unmodified files are found under ``oracle/_ref``,
a byte-for-byte copy made in this container by ``oracle/make_ref.py`` (git-ignored, ships with the tree).
Used by oracle/make_goldens.py to produce tests/golden/*.json and by ``bench.py --impl reference`` /
``bench.py``'s ``cpu_baseline`` leg; never by the product path.
"""
import os
import sys
import types

_LOCAL_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/make_ref.py (git-ignored, ships to the GPU box)


def _default_root() -> str:
    if os.path.isdir("/root/reference/det3d"):
        return "/root/reference"
    return _LOCAL_COPY


REFERENCE_ROOT = os.environ.get("ADSEL_REFERENCE_ROOT") or _default_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "det3d", "selectors", "spatial_temporal_selector.py"))


def _install_shims() -> None:
    if "addict" not in sys.modules:
        addict = types.ModuleType("addict")

        class Dict(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

            def __setattr__(self, k, v):
                self[k] = v

        addict.Dict = Dict
        sys.modules["addict"] = addict
    if "terminaltables" not in sys.modules:
        tt = types.ModuleType("terminaltables")

        class AsciiTable:  # noqa: D401 - import-time placeholder only
            def __init__(self, *a, **k):
                self.table = ""

        tt.AsciiTable = AsciiTable
        sys.modules["terminaltables"] = tt


def load_reference_selectors():
    """Return a dict name -> reference selector class (imported from REFERENCE_ROOT, unmodified)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from det3d import torchie  # noqa: F401  (must import before the selectors package stub)

    if "det3d.selectors" not in sys.modules:
        pkg = types.ModuleType("det3d.selectors")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "det3d", "selectors")]
        sys.modules["det3d.selectors"] = pkg
    from det3d.selectors.spatial_temporal_selector import SpatialTemporalSelector
    from det3d.selectors.spatial_selector import SpatialSelector
    from det3d.selectors.temporal_selector import TemporalSelector
    from det3d.selectors.euclidean_spatial_selector import EuSpatialSelector
    from det3d.selectors.random_selector import RandomSelector

    out = {
        "SpatialTemporalSelector": SpatialTemporalSelector,
        "SpatialSelector": SpatialSelector,
        "TemporalSelector": TemporalSelector,
        "EuSpatialSelector": EuSpatialSelector,
        "RandomSelector": RandomSelector,
    }
    try:
        # feature_selector.py imports ``example_to_device`` from det3d.torchie.apis.train, which pulls in the whole
        # detector stack; it is only used by ``buffer_pred`` (pred=True, needs the trained detector).  A placeholder
        # module lets the UNMODIFIED class be imported and run with pred=False on stored features.
        if "det3d.torchie.apis.train" not in sys.modules:
            apis = sys.modules.get("det3d.torchie.apis") or types.ModuleType("det3d.torchie.apis")
            apis.__path__ = []
            train = types.ModuleType("det3d.torchie.apis.train")

            def example_to_device(*a, **k):   # noqa: D401 - never called with pred=False
                raise RuntimeError("detector forward is out of scope of the oracle")

            train.example_to_device = example_to_device
            apis.train = train
            sys.modules.setdefault("det3d.torchie.apis", apis)
            sys.modules["det3d.torchie.apis.train"] = train
        from det3d.selectors.feature_selector import FeatureSelector
        out["FeatureSelector"] = FeatureSelector
        from det3d.selectors.spatial_feature_selector import SpatialFeatureSelector
        from det3d.selectors.spatial_temporal_feature_selector import SpatialTemporalFeatureSelector
        out["SpatialFeatureSelector"] = SpatialFeatureSelector
        out["SpatialTemporalFeatureSelector"] = SpatialTemporalFeatureSelector
    except Exception:   # pragma: no cover - torch missing etc.; the other selectors stay usable
        pass
    return out
