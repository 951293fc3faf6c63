"""oracle/make_golden_host_api.py -- TEST INFRASTRUCTURE ONLY.

Runs oracle/host_api_cases.py against the UNMODIFIED reference modules (``det3d.utils.registry``,
``det3d.torchie.fileio``; imported from /root/reference with the shims of reference_loader.py) and records the
outcomes -> tests/golden/host_api.json.  tests/test_host.py replays the same table on this package.

    python oracle/make_golden_host_api.py
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import host_api_cases  # noqa: E402
from reference_loader import load_reference_selectors  # noqa: E402


def main():
    load_reference_selectors()                       # installs the shims and puts the reference on sys.path
    from det3d import torchie
    from det3d.utils.registry import Registry, build_from_cfg
    api = {"Registry": Registry, "build_from_cfg": build_from_cfg, "load": torchie.load, "dump": torchie.dump}
    with tempfile.TemporaryDirectory() as wd:
        outcomes = host_api_cases.run_all(api, wd)
    import numpy
    import yaml
    golden = {"source": "det3d/utils/registry.py, det3d/torchie/fileio", "python": sys.version.split()[0],
              "numpy": numpy.__version__, "yaml": yaml.__version__, "outcomes": outcomes}
    out_file = os.path.join(ROOT, "tests", "golden", "host_api.json")
    with open(out_file, "w") as f:
        json.dump(golden, f, indent=1)
    for group, cases in outcomes.items():
        for k, v in cases.items():
            print(group, k, json.dumps(v)[:150])
    print("wrote", out_file)


if __name__ == "__main__":
    main()
