import json
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import adsel_b200  # noqa: E402,F401  (registers the hyphen-named package directory as `adsel_b200`)
from adsel_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    path = os.path.join(GOLDEN_DIR, name)
    if not os.path.exists(path):
        pytest.skip(f"golden fixture {name} not generated")
    with open(path) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def pools(tmp_path_factory):
    """Lazily generated synthetic pools: pools(name) -> dict(infos, logs, paths)."""
    cache = {}

    def get(name):
        if name not in cache:
            spec = getattr(synthetic, name)
            out_dir = str(tmp_path_factory.mktemp(name))
            paths = synthetic.write_pool(spec, out_dir)
            infos, logs = synthetic.generate_pool(spec)
            cache[name] = {"infos": infos, "logs": logs, "paths": paths, "dir": out_dir, "spec": spec}
        return cache[name]

    return get


def first_index_for(n):
    """tools/active_select.py:80 seeds python's RNG with 3407 right before the selector picks."""
    random.seed(3407)
    return random.choice(range(n))


KIND = {"SpatialTemporalSelector": "spatial_temporal", "SpatialSelector": "spatial",
        "TemporalSelector": "temporal", "EuSpatialSelector": "euclid"}
